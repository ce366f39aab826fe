// ekgsim_b200/host/ekg_server.h -- resident evaluator service + client for the ExternalEvaluation adapter.
//
// AMS-DEMO's ExternalEvaluation (ExternalEvaluation.h:95-151) starts `<command line> <homeDir>` once per
// evaluation, from every MPI worker rank (homeDir = process<rank>, GeneralOptimizationAlgorithm.cpp:273-280).
// A process that loads the model, creates a CUDA context and runs the automaton for ONE 1-ms evaluation
// spends > 99 % of its life starting up.  SURVEY 8(f) rank 1 therefore asks for a resident service:
//
//     ekgSim -serve <socket>      holds one Evaluator (model + activation map resident on the GPU), listens on
//                                 a unix-domain socket and answers evaluation requests.  Requests that arrive
//                                 while a batch is on the GPU are evaluated TOGETHER in the next
//                                 Evaluator::evalBatch call, so P worker ranks cost one launch, not P.
//     ekgSim -extern <homeDir>    with EKGSIM_B200_SERVER=<socket> (or -server <socket>): reads
//                                 <homeDir>/input.txt, sends the genes, writes <homeDir>/output.txt.
//
// Wire format (host byte order, same machine): request  = u32 magic 'EKG1' | u32 n_genes | f64 genes[n]
//                                              reply    = i32 status | u32 n | f64 violation | f64 criteria[n]      (status 0)
//                                                         i32 status | u32 n | char message[n]                      (status != 0)
// magic 'EKGQ' asks the server to shut down.
#pragma once

#include <cerrno>
#include <csignal>
#include <cstring>
#include <poll.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#include "ekg_eval.h"

namespace ekg {

constexpr uint32_t kServeMagic = 0x31474B45u;  // "EKG1"
constexpr uint32_t kServeQuit = 0x51474B45u;   // "EKGQ"

inline bool io_all(int fd, void* buf, size_t n, bool writing) {
	char* p = static_cast<char*>(buf);
	while (n) {
		const ssize_t r = writing ? ::send(fd, p, n, MSG_NOSIGNAL) : ::recv(fd, p, n, 0);
		if (r < 0 && errno == EINTR) continue;
		if (r <= 0) return false;
		p += r;
		n -= (size_t)r;
	}
	return true;
}

inline sockaddr_un unix_address(const std::string& path) {
	sockaddr_un a;
	std::memset(&a, 0, sizeof a);
	a.sun_family = AF_UNIX;
	if (path.size() >= sizeof a.sun_path) throw std::runtime_error("socket path too long: " + path);
	std::strncpy(a.sun_path, path.c_str(), sizeof a.sun_path - 1);
	return a;
}

/// client side: one evaluation through a running server; false if no server answers on `path`
inline bool remote_eval(const std::string& path, const std::vector<double>& genes, std::vector<double>& criteria, double& violation) {
	const int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
	if (fd < 0) return false;
	sockaddr_un a = unix_address(path);
	if (::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof a) != 0) { ::close(fd); return false; }
	uint32_t head[2] = {kServeMagic, (uint32_t)genes.size()};
	bool ok = io_all(fd, head, sizeof head, true) && io_all(fd, const_cast<double*>(genes.data()), genes.size() * 8, true);
	int32_t status = 0;
	uint32_t n = 0;
	ok = ok && io_all(fd, &status, 4, false) && io_all(fd, &n, 4, false);
	if (!ok) { ::close(fd); throw std::runtime_error("evaluation server at " + path + " closed the connection"); }
	if (status != 0) {
		if (n > (1u << 20)) n = 1u << 20;
		std::string msg(n, ' ');
		io_all(fd, &msg[0], n, false);
		::close(fd);
		throw std::runtime_error(msg);
	}
	if (n > (1u << 20)) { ::close(fd); throw std::runtime_error("evaluation server at " + path + " sent an implausible reply"); }
	criteria.assign(n, 0.0);
	ok = io_all(fd, &violation, 8, false) && io_all(fd, criteria.data(), (size_t)n * 8, false);
	::close(fd);
	if (!ok) throw std::runtime_error("evaluation server at " + path + " sent a short reply");
	return true;
}

inline void remote_shutdown(const std::string& path) {
	const int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
	if (fd < 0) return;
	sockaddr_un a = unix_address(path);
	if (::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof a) == 0) {
		uint32_t head[2] = {kServeQuit, 0};
		io_all(fd, head, sizeof head, true);
	}
	::close(fd);
}

inline volatile sig_atomic_t& serve_stop_flag() { static volatile sig_atomic_t f = 0; return f; }
inline void serve_on_signal(int) { serve_stop_flag() = 1; }

/// server side; returns when asked to quit or on SIGINT/SIGTERM
inline void serve(Evaluator& ev, const std::string& path, int max_batch = 4096) {
	::unlink(path.c_str());
	const int lfd = ::socket(AF_UNIX, SOCK_STREAM, 0);
	if (lfd < 0) throw std::runtime_error("socket(): " + std::string(std::strerror(errno)));
	sockaddr_un a = unix_address(path);
	if (::bind(lfd, reinterpret_cast<sockaddr*>(&a), sizeof a) != 0 || ::listen(lfd, 1024) != 0) {
		::close(lfd);
		throw std::runtime_error("cannot listen on " + path + ": " + std::strerror(errno));
	}
	std::signal(SIGINT, serve_on_signal);
	std::signal(SIGTERM, serve_on_signal);
	std::cerr << "evaluation server listening on " << path << "\n";

	struct Pending { int fd; std::vector<double> genes; };
	std::vector<int> idle;  // accepted, request not read yet
	size_t served = 0, batches = 0;
	bool quit = false;
	while (!quit && !serve_stop_flag()) {
		std::vector<pollfd> pf(1 + idle.size());
		pf[0].fd = lfd; pf[0].events = POLLIN; pf[0].revents = 0;
		for (size_t i = 0; i < idle.size(); ++i) { pf[1 + i].fd = idle[i]; pf[1 + i].events = POLLIN; pf[1 + i].revents = 0; }
		// block until something arrives; everything that is ready by then forms the next batch
		if (::poll(pf.data(), pf.size(), 500) <= 0) continue;
		if (pf[0].revents & POLLIN) {
			for (;;) {  // take every connection that is already waiting
				pollfd one = {lfd, POLLIN, 0};
				if (::poll(&one, 1, 0) <= 0) break;
				const int fd = ::accept(lfd, nullptr, nullptr);
				if (fd < 0) break;
				// a client that stalls in the middle of a request or does not read its reply must not hold up the others:
				// blocking reads and writes on this connection give up after 2 s and the connection is dropped
				timeval tv;
				tv.tv_sec = 2; tv.tv_usec = 0;
				::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
				::setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
				idle.push_back(fd);
			}
		}
		// requests are tiny (a header and <= a few hundred bytes): a connection that is readable delivers all of it
		std::vector<Pending> batch;
		std::vector<int> still_idle;
		for (int fd : idle) {
			pollfd one = {fd, POLLIN, 0};
			if ((int)batch.size() >= max_batch || ::poll(&one, 1, 0) <= 0) { still_idle.push_back(fd); continue; }
			uint32_t head[2];
			if (!io_all(fd, head, sizeof head, false)) { ::close(fd); continue; }
			if (head[0] == kServeQuit) { quit = true; ::close(fd); continue; }
			if (head[0] != kServeMagic || head[1] > (1u << 20)) { ::close(fd); continue; }
			Pending p;
			p.fd = fd;
			p.genes.assign(head[1], 0.0);
			if (!io_all(fd, p.genes.data(), (size_t)head[1] * 8, false)) { ::close(fd); continue; }
			batch.push_back(std::move(p));
		}
		idle.swap(still_idle);
		if (batch.empty()) continue;

		auto reply_error = [](int fd, const std::string& msg) {
			int32_t status = -1;
			uint32_t n = (uint32_t)msg.size();
			io_all(fd, &status, 4, true) && io_all(fd, &n, 4, true) && io_all(fd, const_cast<char*>(msg.data()), n, true);
			::close(fd);
		};
		// a request with the wrong chromosome length must not take the others down with it
		std::vector<Pending> good;
		for (Pending& p : batch) {
			if (p.genes.size() != ev.numGenes()) {
				std::ostringstream t;
				t << "chromosome size does not agree with the settings of the server (" << p.genes.size() << " != " << ev.numGenes() << ")";
				reply_error(p.fd, t.str());
			} else good.push_back(std::move(p));
		}
		if (good.empty()) continue;
		std::vector<std::vector<double>> sols(good.size()), results;
		std::vector<double> violations;
		for (size_t i = 0; i < good.size(); ++i) sols[i] = good[i].genes;
		try {
			ev.evalBatch(sols, results, violations);
		} catch (std::exception& e) {
			for (Pending& p : good) reply_error(p.fd, e.what());
			continue;
		}
		for (size_t i = 0; i < good.size(); ++i) {
			int32_t status = 0;
			uint32_t n = (uint32_t)results[i].size();
			io_all(good[i].fd, &status, 4, true) && io_all(good[i].fd, &n, 4, true) && io_all(good[i].fd, &violations[i], 8, true) &&
			    io_all(good[i].fd, results[i].data(), (size_t)n * 8, true);
			::close(good[i].fd);
		}
		served += good.size();
		++batches;
	}
	for (int fd : idle) ::close(fd);
	::close(lfd);
	::unlink(path.c_str());
	std::cerr << "evaluation server: " << served << " evaluations in " << batches << " batches\n";
}

}  // namespace ekg
