// tests/cpp/queue_emulation.cpp -- TEST INFRASTRUCTURE: the lock-free work queue of the time-bucket automaton
// (ekgsim_b200/csrc/automaton.cu: ring_publish + the TIMED pop) restated with std::atomic and host threads, one thread
// playing one warp.  K rings of brick ids; a producer reserves a position with one fetch_add on the ring's tail, writes
// the slot and then raises the ring's `published - taken` count; a consumer looks for the earliest ring whose count is
// positive, takes one off the count and only with a positive result claims a position (fetch_add on the head) whose slot
// is filled or about to be; losing the race costs two atomics on the count and no position.  Work items spawn 0..3
// children in buckets [cur, cur + 4]; `pending` (queued + in work) reaching 0 ends the run.  The program checks that every
// item pushed was processed exactly once.  (A first version of the kernel let the losers of a race claim positions and
// abandon them; positions then ran ahead of the tail by a machine width per push, lapped the 16 k-slot rings within
// microseconds and bricks were lost to aliasing -- which is what this restatement was written to pin down.)
//
//     g++ -O2 -std=c++17 -pthread queue_emulation.cpp -o queue_emulation && ./queue_emulation <threads> <items> <ring capacity>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>
#include <chrono>
constexpr int K = 16;
int CAP = 1 << 10;
std::vector<std::atomic<int>> *slots;
std::atomic<unsigned> head[K], tail[K];
std::atomic<int> count[K];
std::atomic<int> cur{0}, pending{0}, processed{0}, pushed{0}, given{0}, dup{0};
std::atomic<long> budget{0};
std::vector<std::atomic<int>> *seen;
static void publish(int bucket, int id) {
	const int r = bucket & (K - 1);
	unsigned pos = tail[r].fetch_add(1);
	(*slots)[(size_t)r * CAP + (pos & (CAP - 1))].store(id);
	count[r].fetch_add(1);
}
static void worker(int w, int nthreads) {
	std::mt19937 rng(w * 7919 + 13);
	unsigned spins = 0;
	for (;;) {
		int b = -1;
		while (b == -1) {
			const int c = cur.load();
			int sel = -1;
			for (int i = 0; i < K; ++i) {
				const int bk = c - 1 + i;
				if (bk < 0) continue;
				const int r = bk & (K - 1);
				if (count[r].load() > 0) { sel = bk; break; }
			}
			if (sel >= 0) {
				const int r = sel & (K - 1);
				if (count[r].fetch_sub(1) <= 0) { count[r].fetch_add(1); given++; continue; }
				unsigned pos = head[r].fetch_add(1);
				std::atomic<int>& s = (*slots)[(size_t)r * CAP + (pos & (CAP - 1))];
				int v;
				while ((v = s.load()) < 0) std::this_thread::yield();
				s.store(-1);
				if (sel > c) { int cc = cur.load(); while (cc < sel && !cur.compare_exchange_weak(cc, sel)) {} }
				b = v;
				continue;
			}
			if (pending.load() == 0) return;
			if (++spins > (1u << 26)) { printf("worker %d gave up, pending %d\n", w, pending.load()); return; }
			std::this_thread::yield();
		}
		// process item b: mark, push 0..3 children into buckets >= cur
		if ((*seen)[b].fetch_add(1) != 0) dup++;
		processed++;
		int n_push = 0;
		int ids[3];
		long left = budget.load();
		if (left > 0) {
			n_push = rng() % 4;
			if (n_push > 3) n_push = 3;
			long got = budget.fetch_sub(n_push);
			if (got < n_push) { n_push = 0; }
		}
		for (int i = 0; i < n_push; ++i) ids[i] = pushed.fetch_add(1);
		if (n_push != 1) pending.fetch_add(n_push - 1);
		if (n_push) {
			const int c = cur.load();
			int key = c + (int)(rng() % 5);
			int bucket = std::min(std::max(key, c), c + K - 2);
			for (int i = 0; i < n_push; ++i) publish(bucket, ids[i]);
		}
	}
}
int main(int argc, char** argv) {
	int nthreads = argc > 1 ? atoi(argv[1]) : 16;
	long total = argc > 2 ? atol(argv[2]) : 2000000;
	CAP = argc > 3 ? atoi(argv[3]) : 1 << 12;
	slots = new std::vector<std::atomic<int>>((size_t)K * CAP);
	for (auto& s : *slots) s.store(-1);
	seen = new std::vector<std::atomic<int>>(total + 64);
	for (auto& s : *seen) s.store(0);
	for (int i = 0; i < K; ++i) { head[i] = 0; tail[i] = 0; count[i] = 0; }
	budget = total;
	// 4 start items in ring 0
	for (int i = 0; i < 4; ++i) { (*slots)[i].store(pushed.fetch_add(1)); }
	tail[0] = 4; count[0] = 4; pending = 4;
	std::vector<std::thread> th;
	auto t0 = std::chrono::steady_clock::now();
	for (int w = 0; w < nthreads; ++w) th.emplace_back(worker, w, nthreads);
	for (auto& t : th) t.join();
	double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	printf("threads %d: pushed %d processed %d pending %d given-up %d duplicates %d cur %d in %.2f s\n", nthreads, pushed.load(), processed.load(), pending.load(), given.load(), dup.load(), cur.load(), s);
	return pushed.load() == processed.load() && pending.load() == 0 ? 0 : 1;
}
