/* oracle/ekg_oracle.c -- TEST INFRASTRUCTURE ONLY (see ekg_oracle.h).
 *
 * Plain-C restatement of the reference hot path.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Pinned against the compiled
 * reference (oracle/_ref) through tests/golden/*: activation map sha256, full-precision ECGs.
 * Compile with -ffp-contract=off: the reference is built without FMA contraction (x86-64
 * baseline, -O2/-O3, no -march), and the automaton must be bit-exact.
 */
#include "ekg_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define START_FLAG 0x1000u /* ShapeElement::layerStartingPoint, matrix.h:90 */

/* ---- Wohlfart.h:195-203 ------------------------------------------------------------------ */
double ekg_oracle_wohlfart_plus(const double k[9], double t) {
	return (1.0 / (1.0 + exp(-k[1] * t)))
	     * (k[2] * ((1.0 - k[3]) * exp(-k[4] * t) + k[3]))
	     * (exp(-k[5] * t) * (1 - pow((1 + exp(-k[7] * (t - k[8])
	            + log(pow(2, (k[7] / k[6])) - 1))), -(k[6] / k[7]))))
	     + k[0];
}

/* ---- simulator.cpp:154-170 ---------------------------------------------------------------- */
double ekg_oracle_ap(const double layer_k[9], double at, double t) {
	double k[9];
	memcpy(k, layer_k, sizeof k);
	k[8] -= at;                       /* ActionPotential::init, simulator.cpp:156 */
	return ekg_oracle_wohlfart_plus(k, t - at); /* operator(), simulator.cpp:168 */
}

/* ---- simulator.h:334-384 ------------------------------------------------------------------ */
int ekg_oracle_neighbourhood(int nbhd, int dif[26 * 3]) {
	int n = 0;
	for (int a0 = 0; a0 < 3; ++a0) for (int a1 = 0; a1 < 3; ++a1) for (int a2 = 0; a2 < 3; ++a2) {
		int d0 = abs(a0 - 1), d1 = abs(a1 - 1), d2 = abs(a2 - 1);
		int keep = 0;
		switch (nbhd) {
		case EKG_ORACLE_NBHD_2D4: { /* callback4N :338-345 */
			int mn = d1 < d2 ? d1 : d2, mx = d1 > d2 ? d1 : d2;
			keep = (mn == 1) && (mx == 1) && (d0 == 0);
			break; }
		case EKG_ORACLE_NBHD_2D8: { /* callback8N :348-354 */
			int mx = d1 > d2 ? d1 : d2;
			keep = (mx == 1) && (d0 == 0);
			break; }
		case EKG_ORACLE_NBHD_3D4: { /* callback3x2N :357-364 -> the 8 cube corners */
			int mn = d0 < d1 ? d0 : d1; if (d2 < mn) mn = d2;
			int mx = d0 > d1 ? d0 : d1; if (d2 > mx) mx = d2;
			keep = (mn == 1) && (mx == 1);
			break; }
		case EKG_ORACLE_NBHD_3D8: { /* callbackCube :367-373 */
			int mx = d0 > d1 ? d0 : d1; if (d2 > mx) mx = d2;
			keep = (mx == 1);
			break; }
		default: return -1;
		}
		if (keep) { dif[3 * n] = a0 - 1; dif[3 * n + 1] = a1 - 1; dif[3 * n + 2] = a2 - 1; ++n; }
	}
	return n;
}

/* ---- binary min-heap on time (std::priority_queue<PriorityQueueEl>, simulator.h:399-413) --- */
typedef struct { double time; int64_t idx; } HeapEl;
typedef struct { HeapEl* a; size_t n, cap; } Heap;

static int heap_push(Heap* h, double time, int64_t idx) {
	if (h->n == h->cap) {
		size_t nc = h->cap ? h->cap * 2 : 1024;
		HeapEl* na = (HeapEl*)realloc(h->a, nc * sizeof(HeapEl));
		if (!na) return -1;
		h->a = na; h->cap = nc;
	}
	size_t i = h->n++;
	while (i > 0) {
		size_t p = (i - 1) / 2;
		if (h->a[p].time <= time) break;
		h->a[i] = h->a[p]; i = p;
	}
	h->a[i].time = time; h->a[i].idx = idx;
	return 0;
}

static HeapEl heap_pop(Heap* h) {
	HeapEl top = h->a[0];
	HeapEl last = h->a[--h->n];
	size_t i = 0;
	for (;;) {
		size_t c = 2 * i + 1;
		if (c >= h->n) break;
		if (c + 1 < h->n && h->a[c + 1].time < h->a[c].time) ++c;
		if (last.time <= h->a[c].time) break;
		h->a[i] = h->a[c]; i = c;
	}
	if (h->n) h->a[i] = last;
	return top;
}

/* ---- simulator.cpp:212-286 ---------------------------------------------------------------- */
int ekg_oracle_activation(const uint16_t* layers_in, int64_t Z, int64_t Y, int64_t X,
                          const double* transfer, int64_t t_rows, int64_t t_cols,
                          double* delay) {
	int64_t n = Z * Y * X;
	int dif[26 * 3];
	/* :251-254: full cube for 3-D shapes, 8-neighbourhood for 2-D ones */
	int nn = ekg_oracle_neighbourhood(Z > 1 ? EKG_ORACLE_NBHD_3D8 : EKG_ORACLE_NBHD_2D8, dif);
	uint16_t* layer = (uint16_t*)malloc((size_t)n * sizeof(uint16_t));
	if (!layer) return -3;
	Heap h = {0, 0, 0};
	int rc = 0;
	(void)t_rows;
	for (int64_t i = 0; i < n; ++i) {
		delay[i] = 0.0;
		layer[i] = layers_in[i];
		if (layer[i] & START_FLAG) {            /* :261-264 */
			heap_push(&h, 1.0, i);
			layer[i] = (uint16_t)(layer[i] - START_FLAG);
		}
	}
	if (h.n == 0) { rc = -1; goto done; }       /* :274-277 */
	while (h.n) {                               /* exciteElement :212-246 */
		HeapEl act = heap_pop(&h);
		if (delay[act.idx] != 0) continue;      /* :219 */
		delay[act.idx] = act.time;              /* :221 */
		int64_t z = act.idx / (Y * X), y = (act.idx / X) % Y, x = act.idx % X;
		uint16_t l = layer[act.idx];
		for (int k = 0; k < nn; ++k) {
			int64_t z2 = z - dif[3 * k], y2 = y - dif[3 * k + 1], x2 = x - dif[3 * k + 2]; /* :230 */
			if (z2 < 0 || z2 >= Z || y2 < 0 || y2 >= Y || x2 < 0 || x2 >= X) continue;
			int64_t j = (z2 * Y + y2) * X + x2;
			if (layer[j] == 0 || delay[j] != 0) continue;                                   /* :232 */
			if (l >= t_cols || layer[j] >= t_cols) { rc = -2; goto done; }                  /* :234-238 */
			double lag = transfer[(int64_t)l * t_cols + layer[j]];                          /* :239 */
			int sq = dif[3 * k] * dif[3 * k] + dif[3 * k + 1] * dif[3 * k + 1] + dif[3 * k + 2] * dif[3 * k + 2];
			lag *= sqrt((double)sq);                                                        /* :240 */
			if (heap_push(&h, act.time + lag, j)) { rc = -3; goto done; }                   /* :241 */
		}
	}
done:
	free(h.a);
	free(layer);
	return rc;
}

/* ---- simulator.cpp:561-621 ---------------------------------------------------------------- */
static uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static uint64_t mix64(uint64_t x) {
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
	return x;
}

int64_t ekg_oracle_ap_classes(const uint16_t* layers, const double* delay,
                              int64_t n, int64_t n_layers, int64_t* ap_index) {
	/* std::map<double,size_t> per layer keyed by the exact delay value (:566-590); here one
	 * open-addressing table keyed by (layer, delay bits); -0.0 cannot occur (delays are 0 or >= 1) */
	size_t cap = 1024;
	int64_t occupied = 0;
	for (int64_t i = 0; i < n; ++i) occupied += (layers[i] & ~START_FLAG) != 0;
	while (cap < (size_t)occupied * 2 + 16) cap <<= 1;
	uint64_t* keyb = (uint64_t*)malloc(cap * 8);
	uint16_t* keyl = (uint16_t*)calloc(cap, 2);
	int64_t* val = (int64_t*)malloc(cap * 8);
	if (!keyb || !keyl || !val) { free(keyb); free(keyl); free(val); return -3; }
	int64_t next = 0;
	(void)n_layers;
	for (int64_t i = 0; i < n; ++i) {
		uint16_t l = (uint16_t)(layers[i] & ~START_FLAG);
		if (l == 0) { ap_index[i] = -1; continue; }   /* :576-577 */
		uint64_t b = dbits(delay[i]);
		size_t hsh = (size_t)(mix64(b ^ ((uint64_t)l << 52)) & (cap - 1));
		for (;;) {
			if (keyl[hsh] == 0) { keyl[hsh] = l; keyb[hsh] = b; val[hsh] = next; ap_index[i] = next++; break; }
			if (keyl[hsh] == l && keyb[hsh] == b) { ap_index[i] = val[hsh]; break; }
			hsh = (hsh + 1) & (cap - 1);
		}
	}
	free(keyb); free(keyl); free(val);
	return next;
}

/* ---- simulator.cpp:452-550 ---------------------------------------------------------------- */
int64_t ekg_oracle_run_direct(const uint16_t* layers, const double* delay,
                              int64_t Z, int64_t Y, int64_t X,
                              const double* layer_k, int64_t n_layers,
                              const double* leads, int64_t n_leads, int nbhd,
                              double t_start, double t_step, double total_time,
                              double* ecg) {
	int dif[26 * 3];
	int nn = ekg_oracle_neighbourhood(nbhd, dif);
	if (nn < 0) return -1;
	int64_t steps = (int64_t)ceil(total_time / t_step);          /* :471 */
	double* cell = (double*)malloc((size_t)steps * 8);
	double* dip = (double*)malloc((size_t)steps * 3 * 8);
	if (!cell || !dip) { free(cell); free(dip); return -3; }
	for (int64_t i = 0; i < n_leads * steps; ++i) ecg[i] = 0.0;   /* :480-481 */
	double end_time = total_time + t_start;                      /* :483 */

	for (int64_t z = 0; z < Z; ++z) for (int64_t y = 0; y < Y; ++y) for (int64_t x = 0; x < X; ++x) { /* :489 */
		int64_t c = (z * Y + y) * X + x;
		uint16_t l = (uint16_t)(layers[c] & ~START_FLAG);
		if (l == 0) continue;                                    /* :492 */
		if (l > n_layers) { free(cell); free(dip); return -2; }
		double kc[9];
		memcpy(kc, layer_k + 9 * (l - 1), sizeof kc);
		kc[8] -= delay[c];                                       /* setApIndices -> init, :594, :156 */
		double sim_time = t_start;
		for (int64_t i = 0; i < steps; ++i) {                    /* :496-500 */
			cell[i] = ekg_oracle_wohlfart_plus(kc, sim_time - delay[c]);
			sim_time += t_step;
		}
		memset(dip, 0, (size_t)steps * 3 * 8);                   /* :506-507 */
		for (int k = 0; k < nn; ++k) {                           /* :511 */
			int64_t z2 = z - dif[3 * k], y2 = y - dif[3 * k + 1], x2 = x - dif[3 * k + 2]; /* :514 */
			if (z2 < 0 || z2 >= Z || y2 < 0 || y2 >= Y || x2 < 0 || x2 >= X) continue; /* zero border :458-463 */
			int64_t j = (z2 * Y + y2) * X + x2;
			uint16_t l2 = (uint16_t)(layers[j] & ~START_FLAG);
			if (l2 == 0) continue;                               /* :515 */
			if (l2 > n_layers) { free(cell); free(dip); return -2; }
			double kn[9];
			memcpy(kn, layer_k + 9 * (l2 - 1), sizeof kn);
			kn[8] -= delay[j];
			int64_t next = 0;
			for (double st = t_start; st < end_time && next < steps; st += t_step) { /* :519 */
				double vd = ekg_oracle_wohlfart_plus(kn, st - delay[j]) - cell[next];   /* :520 */
				dip[3 * next]     += dif[3 * k] * vd;                                   /* :521 */
				dip[3 * next + 1] += dif[3 * k + 1] * vd;
				dip[3 * next + 2] += dif[3 * k + 2] * vd;
				++next;
			}
		}
		for (int64_t m = 0; m < n_leads; ++m) {                  /* :530-537 */
			/* position of the voxel = its index in the zero-bordered matrix: (z+1,y+1,x+1) :487,:531 */
			double p0 = leads[3 * m] - (double)(z + 1);
			double p1 = leads[3 * m + 1] - (double)(y + 1);
			double p2 = leads[3 * m + 2] - (double)(x + 1);
			double sq = p0 * p0; sq += p1 * p1; sq += p2 * p2;   /* sqrLength, Hypermatrix.h:250-256 */
			double inv = 1.0 / (sq * sqrt(sq));                  /* pow3Length :258-261, :533 */
			p0 *= inv; p1 *= inv; p2 *= inv;
			double* e = ecg + m * steps;
			for (int64_t t = 0; t < steps; ++t) {
				double s = 0; s += p0 * dip[3 * t]; s += p1 * dip[3 * t + 1]; s += p2 * dip[3 * t + 2]; /* :540-546 */
				e[t] += s;                                       /* :535 */
			}
		}
	}
	free(cell); free(dip);
	return steps;
}

/* ---- class-factored restatement (same sum, re-associated) --------------------------------- */
int64_t ekg_oracle_run_factored(const uint16_t* layers, const double* delay,
                                int64_t Z, int64_t Y, int64_t X,
                                const double* layer_k, int64_t n_layers,
                                const double* leads, int64_t n_leads, int nbhd,
                                double t_start, double t_step, double total_time,
                                double* ecg) {
	return ekg_oracle_run_factored_slab(layers, delay, Z, Y, X, layer_k, n_layers, leads, n_leads, nbhd,
	                                    t_start, t_step, total_time, 0, Z, ecg);
}

/* Partial ECG of the z-slab [z0, z1): only the terms G_c V_c(t) of voxels c inside the slab, with
 * G_c built from the WHOLE model's neighbourhood.  Slabs partition the voxels, so the partial ECGs
 * of a partition of [0, Z) add up to the full ECG (multi-GPU sharding of one model). */
int64_t ekg_oracle_run_factored_slab(const uint16_t* layers, const double* delay,
                                     int64_t Z, int64_t Y, int64_t X,
                                     const double* layer_k, int64_t n_layers,
                                     const double* leads, int64_t n_leads, int nbhd,
                                     double t_start, double t_step, double total_time,
                                     int64_t z0, int64_t z1, double* ecg) {
	int dif[26 * 3];
	int nn = ekg_oracle_neighbourhood(nbhd, dif);
	if (nn < 0) return -1;
	int64_t n = Z * Y * X;
	int64_t steps = (int64_t)ceil(total_time / t_step);
	int64_t* cls = (int64_t*)malloc((size_t)n * 8);
	if (!cls) return -3;
	int64_t K = ekg_oracle_ap_classes(layers, delay, n, n_layers, cls);
	if (K < 0) { free(cls); return -3; }
	double* G = (double*)calloc((size_t)(K * n_leads), 8);        /* G[class][lead] */
	double* ck = (double*)malloc((size_t)K * 10 * 8);             /* class: 9 coeffs (k8 shifted) + at */
	double* w = (double*)malloc((size_t)n_leads * 3 * 8);
	if (!G || !ck || !w) { free(cls); free(G); free(ck); free(w); return -3; }

	for (int64_t z = 0; z < Z; ++z) for (int64_t y = 0; y < Y; ++y) for (int64_t x = 0; x < X; ++x) {
		int64_t c = (z * Y + y) * X + x;
		uint16_t l = (uint16_t)(layers[c] & ~START_FLAG);
		if (l == 0) continue;
		if (l > n_layers) { free(cls); free(G); free(ck); free(w); return -2; }
		double* kk = ck + 10 * cls[c];
		memcpy(kk, layer_k + 9 * (l - 1), 9 * 8);
		kk[8] -= delay[c];
		kk[9] = delay[c];
		for (int64_t m = 0; m < n_leads; ++m) {
			double p0 = leads[3 * m] - (double)(z + 1), p1 = leads[3 * m + 1] - (double)(y + 1), p2 = leads[3 * m + 2] - (double)(x + 1);
			double sq = p0 * p0; sq += p1 * p1; sq += p2 * p2;
			double inv = 1.0 / (sq * sqrt(sq));
			w[3 * m] = p0 * inv; w[3 * m + 1] = p1 * inv; w[3 * m + 2] = p2 * inv;
		}
		/* ECG_l(t) += sum_dif (w_c . dif) (V_{c-dif}(t) - V_c(t))  ->  scatter the coefficient */
		for (int k = 0; k < nn; ++k) {
			int64_t z2 = z - dif[3 * k], y2 = y - dif[3 * k + 1], x2 = x - dif[3 * k + 2];
			if (z2 < 0 || z2 >= Z || y2 < 0 || y2 >= Y || x2 < 0 || x2 >= X) continue;
			int64_t j = (z2 * Y + y2) * X + x2;
			if ((layers[j] & ~START_FLAG) == 0) continue;
			for (int64_t m = 0; m < n_leads; ++m) {
				double wd = w[3 * m] * dif[3 * k] + w[3 * m + 1] * dif[3 * k + 1] + w[3 * m + 2] * dif[3 * k + 2];
				if (z2 >= z0 && z2 < z1) G[cls[j] * n_leads + m] += wd;
				if (z >= z0 && z < z1) G[cls[c] * n_leads + m] -= wd;
			}
		}
	}
	for (int64_t i = 0; i < n_leads * steps; ++i) ecg[i] = 0.0;
	double* v = (double*)malloc((size_t)steps * 8);
	if (!v) { free(cls); free(G); free(ck); free(w); return -3; }
	for (int64_t k = 0; k < K; ++k) {
		const double* kk = ck + 10 * k;
		double st = t_start;
		for (int64_t t = 0; t < steps; ++t) { v[t] = ekg_oracle_wohlfart_plus(kk, st - kk[9]); st += t_step; }
		for (int64_t m = 0; m < n_leads; ++m) {
			double g = G[k * n_leads + m];
			double* e = ecg + m * steps;
			for (int64_t t = 0; t < steps; ++t) e[t] += g * v[t];
		}
	}
	free(v); free(cls); free(G); free(ck); free(w);
	return steps;
}

/* ---- simulator.cpp:552-559 ---------------------------------------------------------------- */
int64_t ekg_oracle_run_approximation(const double* layer_k, int64_t n_layers,
                                     double t_start, double t_step, double total_time,
                                     double delay, double* out) {
	int64_t n = (int64_t)(total_time / t_step);               /* :553 */
	for (int64_t i = 0; i < n; ++i) {
		double st = t_start + i * t_step;                     /* :556 */
		out[i] = ekg_oracle_wohlfart_plus(layer_k, st + delay)
		       - ekg_oracle_wohlfart_plus(layer_k + 9 * (n_layers - 1), st); /* :557 (layer APs have at = 0) */
	}
	return n;
}

/* ==== layer-AP construction (evaluation glue) ================================================
 * Plain restatement of the per-layer fit: every price evaluation is a full AP evaluation, as
 * in the reference (no factor caching). */

/* ---- Wohlfart.h:206-223 (WohlfartPlus::apd90) + interpolate :226-229 -------------------- */
double ekg_oracle_apd90(const double k[9]) {
	double d[1000];
	for (int i = 0; i < 1000; ++i) d[i] = ekg_oracle_wohlfart_plus(k, (double)i);   /* :210-212 */
	double target = k[0] + k[2] * 0.1;                                             /* :214 */
	double time = -1.0;
	for (int i = 1; i < 1000; ++i)
		if ((d[i - 1] > target) && (d[i] <= target)) {                              /* :217 */
			double a = (double)(i - 1), wa = d[i - 1] - target, b = (double)i, wb = target - d[i];
			time = (a * wa + b * wb) / (wa + wb);                                  /* :228 */
		}
	return time;
}

typedef struct { double k, n, x1, x2; } fit_connector;   /* LineConnector, sim.cpp:95-109 */

static size_t fit_cap700(double apd) {   /* std::min((size_t)700, (size_t)ceil(apd)), sim.cpp:206,208 */
	double c = ceil(apd);
	if (c < 0) return 700;               /* apd90 = -1: the size_t cast wraps to a huge value */
	return c > 700.0 ? 700 : (size_t)c;
}

/* ---- sim.cpp:178-245 (distributeConnectors) --------------------------------------------- */
static void fit_distribute_connectors(const double ap1[9], const double ap2[9], fit_connector out[15]) {
	const size_t numPoints = 15, startX = 10;                                      /* :185, :192 */
	const double xScaleFactor = 0.1;                                               /* :199 */
	double apd1 = ekg_oracle_apd90(ap1), apd2 = ekg_oracle_apd90(ap2);             /* :188-189 */
	/* the reference holds 700 samples and reads element [700] when a walk reaches the cap (:224, :228);
	 * here that element exists */
	double y1[701], y2[701];
	for (size_t i = 0; i < 701; ++i) { y1[i] = ekg_oracle_wohlfart_plus(ap1, (double)i); y2[i] = ekg_oracle_wohlfart_plus(ap2, (double)i); }
	double len1 = 0.0, len2 = 0.0;
	for (size_t i = startX + 1; i < fit_cap700(apd1); ++i) len1 += sqrt((y1[i] - y1[i - 1]) * (y1[i] - y1[i - 1]) + xScaleFactor); /* :206-207 */
	for (size_t i = startX; i < fit_cap700(apd2); ++i) len2 += sqrt((y2[i] - y2[i - 1]) * (y2[i] - y2[i - 1]) + xScaleFactor);     /* :208-209 */
	double currentLen1 = 0.0, currentLen2 = 0.0;
	size_t i1 = startX, i2 = startX;
	for (size_t i = 0; i < numPoints; ++i) {                                       /* :218-244 */
		double targetLen = i * len1 / (numPoints - 2);
		for (double l1 = currentLen1; (l1 < targetLen) && (i1 < 700); ++i1) l1 += sqrt((y1[i1 + 1] - y1[i1]) * (y1[i1 + 1] - y1[i1]) + xScaleFactor);
		currentLen1 = targetLen;
		targetLen = i * len2 / (numPoints - 2);
		for (double l2 = currentLen2; (l2 < targetLen) && (i2 < 700); ++i2) l2 += sqrt((y2[i2 + 1] - y2[i2]) * (y2[i2 + 1] - y2[i2]) + xScaleFactor);
		currentLen2 = targetLen;
		fit_connector lc;
		lc.x1 = (double)i1;
		lc.x2 = (double)i2;
		if (i1 == i2) lc.x2 += 0.001;                                              /* :233-234 */
		lc.k = (y1[i1] - y2[i2]) / (lc.x1 - lc.x2);                                /* :236 */
		lc.n = y2[i2] - lc.k * i2;                                                 /* :238 */
		out[i] = lc;
	}
}

/* ---- sim.cpp:249-257 (setupRatio) ------------------------------------------------------- */
static void fit_setup_ratio(const fit_connector c[15], double ratio, double pts[30]) {
	for (int i = 0; i < 15; ++i) {
		pts[2 * i] = c[i].x1 + (c[i].x2 - c[i].x1) * ratio;                        /* getX :107-109 */
		pts[2 * i + 1] = c[i].k * pts[2 * i] + c[i].n;                             /* operator() :102-104 */
	}
}

/* ---- sim.cpp:157-168 (price: sum of squared misses; `derivatives` is always empty) ------ */
static double fit_price(const double k[9], const double pts[30]) {
	double sum = 0;
	for (int i = 0; i < 30; i += 2) {
		double e = ekg_oracle_wohlfart_plus(k, pts[i]) - pts[i + 1];
		sum += e * e;
	}
	return sum;
}

/* ---- nonlinearFit.h:92-168 (steepestDescend) -------------------------------------------- */
static int fit_steepest_descend(const double pts[30], double x0[9], const double d[9], double stepSize, double epsilon, int iterations) {
	double grad[9], oldGrad[9], move[9], x1[9];
	for (int i = 0; i < 9; ++i) { grad[i] = x0[i]; oldGrad[i] = 0; move[i] = d[i]; }
	double y0 = fit_price(x0, pts);                                                /* :106 */
	for (; (iterations > 0) && (y0 > epsilon); --iterations) {                     /* :116 */
		int stepChange = 0;
		for (int i = 0; i < 9; ++i) {
			memcpy(x1, x0, sizeof x1);
			if (d[i] != 0) {
				x1[i] += d[i] * .001;                                              /* :125 */
				grad[i] = (fit_price(x1, pts) - y0) / (d[i] * .001);               /* :126 */
				if (grad[i] * oldGrad[i] < 0) { move[i] *= 0.5; stepChange = 1; }  /* :128-131 */
				else if (fabs(grad[i]) > 0.75 * fabs(oldGrad[i])) move[i] *= 1.5;  /* :132-134 */
			} else grad[i] = 0;
		}
		memcpy(oldGrad, grad, sizeof grad);                                        /* :144 */
		memcpy(x1, x0, sizeof x1);
		for (int i = 0; i < 9; ++i) x1[i] -= stepSize * ((grad[i] > 0) ? move[i] : -move[i]); /* :148-150 */
		double y1 = fit_price(x1, pts);
		if (y1 < y0) { y0 = y1; memcpy(x0, x1, sizeof x1); }                       /* :153-155 */
		else if (!stepChange) stepSize *= 0.5;                                     /* :156-159 */
	}
	return iterations;
}

/* ---- sim.cpp:751-821 (n_border = 2) and :825-916 (n_border = 3) --------------------------
 * border_k: [n_border][9]; out: [n_layers][9].  The gene unpacking / violation part of those
 * functions is not restated here (it is a copy of genes into coefficients). */
int ekg_oracle_fit_layers(const double* border_k, int64_t n_border, int64_t n_layers, int64_t mid,
                          const double d9[9], double step, double eps, int64_t iterations, double* out) {
	fit_connector conn[15];
	double pts[30];
	if (n_layers < n_border || (n_border != 2 && n_border != 3)) return -1;
	const double* front = border_k;
	const double* back = border_k + 9 * (n_border - 1);
	memcpy(out, front, 72);
	if (n_border == 3) memcpy(out + 9 * mid, border_k + 9, 72);                    /* :856-865 */
	memcpy(out + 9 * (n_layers - 1), back, 72);
	if (n_border == 2) {
		fit_distribute_connectors(out, out + 9 * (n_layers - 1), conn);            /* :793-794 */
		for (int64_t i = 1; i < n_layers - 1; ++i) {
			double ratio = i / (double)(n_layers - 1);                             /* :802 */
			fit_setup_ratio(conn, ratio, pts);
			for (int q = 0; q < 9; ++q) out[9 * i + q] = front[q] * (1 - ratio) + back[q] * ratio;   /* combineAps :705-709 */
			fit_steepest_descend(pts, out + 9 * i, d9, step, eps, (int)iterations);          /* :806 */
		}
		return 0;
	}
	const double* midk = out + 9 * mid;
	for (int64_t i = 1; i < n_layers - 1; ++i) {                                   /* :883-902 */
		double ratio;
		if (i < mid) {
			ratio = i / (double)mid;
			if (i == 1) fit_distribute_connectors(out, midk, conn);                /* :887 (same connectors for every i < mid) */
			fit_setup_ratio(conn, ratio, pts);
			for (int q = 0; q < 9; ++q) out[9 * i + q] = out[q] * (1 - ratio) + midk[q] * ratio;
		} else if (i > mid) {
			ratio = (i - mid) / (double)(n_layers - mid - 1);
			if (i - mid == 1) fit_distribute_connectors(midk, out + 9 * (n_layers - 1), conn);   /* :894-895 */
			fit_setup_ratio(conn, ratio, pts);
			for (int q = 0; q < 9; ++q) out[9 * i + q] = midk[q] * (1 - ratio) + out[9 * (n_layers - 1) + q] * ratio;
		} else continue;
		fit_steepest_descend(pts, out + 9 * i, d9, step, eps, (int)iterations);    /* :900-901 */
	}
	return 0;
}
