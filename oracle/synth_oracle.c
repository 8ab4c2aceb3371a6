/*
 * TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT. See synth_oracle.h for role and parity status (PINNED).
 *
 * Plain-C restatement of the reference's synthesis stage. Every function names the reference lines it follows
 * (paths relative to /root/reference). Compiled with -ffp-contract=off so that a*b+c is two roundings, like the
 * reference's scalar C++ build (g++ -O2 on x86-64 does not contract either).
 */
#define _USE_MATH_DEFINES
#include "synth_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

/* ------------------------------------------------------------------------------------------------------------
 * floor1_inverse_dB_table — src/inverse_db_table.h:13-78 (Vorbis I spec 10.1).
 * The spec table is fromdB((i-255)*0.546875) with fromdB(x)=exp(x*0.11512925), printed with 8 significant
 * digits; the reference's float literals are those decimals. Regenerating it the same way reproduces all 256
 * floats bit-exactly (checked against the reference header by tests/golden/make_golden.py).
 * ---------------------------------------------------------------------------------------------------------- */
static float g_inv_db[256];
static int g_inv_db_ready = 0;

const float* por_inverse_db_table(void) {
	if(!g_inv_db_ready) {
		for(int i = 0; i < 256; ++i) {
			char buf[64];
			double v = exp((double)(i - 255) * 0.546875 * 0.11512925);
			snprintf(buf, sizeof buf, "%.7e", v);
			g_inv_db[i] = strtof(buf, NULL);
		}
		g_inv_db_ready = 1;
	}
	return g_inv_db;
}

/* ------------------------------------------------------------------------------------------------------------
 * floor1
 * ---------------------------------------------------------------------------------------------------------- */

/* src/Utils.hpp:60-118. low: the earlier post with the greatest x below xs[i]; high: the earlier post with the
 * smallest x above xs[i]. Ties cannot occur among distinct xs; on equal candidates the reference keeps the
 * first one found (strict comparisons), which the scan below reproduces. -1 = none. */
void por_floor1_neighbors(const uint16_t* xs, int posts, int* lo, int* hi) {
	for(int i = 0; i < posts; ++i) {
		int best_lo = -1, best_hi = -1;
		for(int j = 0; j < i; ++j) {
			if(xs[j] < xs[i] && (best_lo < 0 || xs[j] > xs[best_lo])) best_lo = j;
			if(xs[j] > xs[i] && (best_hi < 0 || xs[j] < xs[best_hi])) best_hi = j;
		}
		lo[i] = best_lo;
		hi[i] = best_hi;
	}
}

/* src/ParseOggVorbis.hpp:458-469 (std::sort by x; xs are distinct in valid setups, so the order is unique) */
void por_floor1_sort(const uint16_t* xs, int posts, int* sorted_idx) {
	for(int i = 0; i < posts; ++i) sorted_idx[i] = i;
	for(int i = 1; i < posts; ++i) { /* insertion sort, stable */
		int v = sorted_idx[i], j = i;
		while(j > 0 && xs[sorted_idx[j - 1]] > xs[v]) { sorted_idx[j] = sorted_idx[j - 1]; --j; }
		sorted_idx[j] = v;
	}
}

/* src/Utils.hpp:122-137, unsigned 32-bit arithmetic as in the reference (T = uint32_t) */
static uint32_t point_on_line(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t X) {
	uint32_t adx = x1 - x0;
	int up = y1 >= y0;
	uint32_t ady = up ? (y1 - y0) : (y0 - y1);
	uint32_t off = (ady * (X - x0)) / adx;
	return up ? y0 + off : y0 - off;
}

static uint32_t floor1_range(int multiplier) { /* src/ParseOggVorbis.hpp:486-492 */
	static const uint32_t r[4] = {256, 128, 86, 64};
	return r[(multiplier - 1) & 3];
}

/* src/ParseOggVorbis.hpp:521-559 */
uint32_t por_floor1_unwrap(const uint16_t* xs, int posts, int multiplier, const uint32_t* ys,
                           uint32_t* final_ys, uint8_t* step2_flag) {
	int lo[POV_MAX_POSTS], hi[POV_MAX_POSTS];
	uint32_t status = 0;
	const uint32_t range = floor1_range(multiplier);
	por_floor1_neighbors(xs, posts, lo, hi);
	memset(step2_flag, 0, (size_t) posts);
	step2_flag[0] = 1; step2_flag[1] = 1;
	final_ys[0] = ys[0]; final_ys[1] = ys[1];
	for(int i = 2; i < posts; ++i) {
		int l = lo[i], h = hi[i];
		uint32_t predicted = point_on_line(xs[l], final_ys[l], xs[h], final_ys[h], xs[i]);
		uint32_t val = ys[i];
		if(predicted > range) status |= POV_PKT_FLOOR_PREDICTED; /* hpp:536, fatal in the reference */
		uint32_t high_room = range - predicted, low_room = predicted;
		uint32_t room = (high_room < low_room ? high_room : low_room) * 2;
		if(val == 0) {
			final_ys[i] = predicted;
		} else {
			step2_flag[l] = 1; step2_flag[h] = 1; step2_flag[i] = 1;
			if(val >= room)
				final_ys[i] = (high_room > low_room) ? val - low_room + predicted
				                                     : predicted - val + high_room - 1;
			else
				final_ys[i] = (val & 1) ? predicted - (val + 1) / 2 : predicted + val / 2;
		}
	}
	return status;
}

/* src/Utils.hpp:143-183: incremental (error-accumulating) integer line, clipped to [0,n) */
static void raster_line(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t n, uint32_t* v) {
	if(x0 >= n) return;
	uint32_t dx = x1 - x0;
	int up = y1 >= y0;
	uint32_t dy = up ? y1 - y0 : y0 - y1;
	uint32_t base = dy / dx, rem = dy - base * dx, err = 0, y = y0;
	v[x0] = y0;
	for(uint32_t x = x0 + 1; x < x1 && x < n; ++x) {
		uint32_t step = base;
		err += rem;
		if(err >= dx) { err -= dx; step += 1; }
		y = up ? y + step : y - step;
		v[x] = y;
	}
}

/* src/ParseOggVorbis.hpp:563-585 */
void por_floor1_render(const uint16_t* xs, int posts, int multiplier, const uint32_t* final_ys,
                       const uint8_t* step2_flag, uint32_t n, uint32_t* floor_out) {
	int order[POV_MAX_POSTS];
	por_floor1_sort(xs, posts, order);
	memset(floor_out, 0, (size_t) n * sizeof(uint32_t));
	uint32_t lx = 0, hx = 0, ly = final_ys[order[0]] * (uint32_t) multiplier, hy = 0;
	for(int s = 1; s < posts; ++s) {
		int i = order[s];
		if(!step2_flag[i]) continue;
		hx = xs[i];
		hy = final_ys[i] * (uint32_t) multiplier;
		raster_line(lx, ly, hx, hy, n, floor_out);
		lx = hx; ly = hy;
	}
	if(hx < n) raster_line(hx, hy, n, hy, n, floor_out);
}

/* src/ParseOggVorbis.hpp:586-589 */
uint32_t por_floor1_db(const uint32_t* floor, uint32_t n, float* out) {
	const float* t = por_inverse_db_table();
	uint32_t status = 0;
	for(uint32_t i = 0; i < n; ++i) {
		if(floor[i] < 256) out[i] = t[floor[i]];
		else status |= POV_PKT_FLOOR_RANGE;
	}
	return status;
}

/* ------------------------------------------------------------------------------------------------------------
 * residue application — src/ParseOggVorbis.hpp:696-757 (types 0/1; type 2 = type 1 on one interleaved vector,
 * hpp:685-694, handled by the caller). Entropy decode already happened on the host; this replays the adds in
 * the reference's order: pass, partition, channel, vector.
 * ---------------------------------------------------------------------------------------------------------- */
uint32_t por_residue_apply(const pov_residue* res, const pov_codebook* books, uint32_t n_books,
                           uint32_t nch, const uint8_t* ch_used, uint32_t decode_len, int entry_bits,
                           const uint8_t* payload, uint64_t* consumed, float* out) {
	uint32_t status = 0;
	uint32_t lb = res->begin < decode_len ? res->begin : decode_len;     /* hpp:697 */
	uint32_t le = res->end < decode_len ? res->end : decode_len;         /* hpp:698 */
	uint32_t parts = (le - lb) / res->partition_size;                    /* hpp:706 */
	uint32_t n_entries;
	memcpy(&n_entries, payload, 4);
	const uint8_t* cls = payload + 4;
	uint64_t cls_bytes = ((uint64_t) nch * parts + 3) & ~(uint64_t) 3;
	const uint8_t* ent = cls + cls_bytes;
	uint64_t ent_bytes = ((uint64_t) n_entries * (entry_bits / 8) + 3) & ~(uint64_t) 3;
	if(consumed) *consumed = 4 + cls_bytes + ent_bytes;
	uint32_t cur = 0;
	for(uint32_t pass = 0; pass < 8; ++pass) {
		for(uint32_t p = 0; p < parts; ++p) {
			for(uint32_t j = 0; j < nch; ++j) {
				if(!ch_used[j]) continue;                                        /* hpp:727 */
				uint32_t c = cls[(uint64_t) j * parts + p];
				uint32_t book = res->books[c * 8 + pass];                        /* hpp:729 */
				if(book == POV_NO_BOOK) continue;
				if(book >= n_books) { status |= POV_PKT_VQ_ENTRY; continue; }
				const pov_codebook* cb = &books[book];
				float* v = out + (uint64_t) j * decode_len;
				uint32_t offset = lb + p * res->partition_size;                 /* hpp:733 */
				uint32_t dim = cb->dim;
				uint32_t nvec = res->partition_size / dim;
				for(uint32_t k = 0; k < nvec; ++k) {
					if(cur >= n_entries) { status |= POV_PKT_VQ_ENTRY; break; }
					uint32_t e;
					if(entry_bits == 16) { uint16_t t; memcpy(&t, ent + 2ull * cur, 2); e = t; }
					else memcpy(&e, ent + 4ull * cur, 4);
					++cur;
					if(cb->lookup_type == 0 || e >= cb->n_entries || !cb->vq) { /* hpp:369-370 */
						status |= POV_PKT_VQ_ENTRY;
						continue;
					}
					const float* vec = cb->vq + (uint64_t) e * dim;
					if(res->type == 0) {                                         /* hpp:736-742 */
						for(uint32_t l = 0; l < dim; ++l) v[offset + k + l * nvec] += vec[l];
					} else {                                                     /* hpp:746-751 */
						for(uint32_t l = 0; l < dim; ++l) v[offset + k * dim + l] += vec[l];
					}
				}
			}
		}
	}
	return status;
}

/* src/ParseOggVorbis.hpp:1219-1240 */
void por_inverse_coupling(float* mag, float* ang, uint32_t n) {
	for(uint32_t i = 0; i < n; ++i) {
		float m = mag[i], a = ang[i], m2 = m, a2 = a;
		if(m > 0) {
			if(a > 0) a2 = m - a;
			else { a2 = m; m2 = m + a; }
		} else {
			if(a > 0) a2 = m + a;
			else { a2 = m; m2 = m - a; }
		}
		mag[i] = m2; ang[i] = a2;
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * inverse MDCT. The reference's src/mdct.cpp is libvorbis' split-radix code; north_star forbids porting it.
 * Contract (probed against src/mdct.cpp:433-527 through its extern "C" mdct_backward):
 *     y[m] = sum_{k<n/2} X[k] * cos( (2*pi/n) * (m + 1/2 + n/4) * (k + 1/2) ),  m < n,  no scaling.
 * ---------------------------------------------------------------------------------------------------------- */
void por_imdct_closed(const float* in, uint32_t n, float* out) {
	const uint32_t M = n / 2;
	/* cos table over the 4n-periodic integer phase (2m+1+n/2)(2k+1):  angle = pi/(2n) * phase */
	double* ct = (double*) malloc(sizeof(double) * 4 * (size_t) n);
	for(uint32_t i = 0; i < 4 * n; ++i) ct[i] = cos(M_PI / (2.0 * n) * (double) i);
	for(uint32_t m = 0; m < n; ++m) {
		double acc = 0;
		uint64_t a = 2ull * m + 1 + n / 2;
		for(uint32_t k = 0; k < M; ++k) acc += (double) in[k] * ct[(a * (2ull * k + 1)) % (4ull * n)];
		out[m] = (float) acc;
	}
	free(ct);
}

/* Fast path (our own algorithm, the same decomposition the CUDA kernels use — DESIGN.md §IMDCT):
 *   DCT-IV of size M=n/2 through one complex FFT of size M/2:
 *     t[j] = (X[2j] + i X[M-1-2j]) * w[j],  w[j] = exp(-i*pi*(8j+1)/(8M));  T = FFT(t);  c = T * w
 *     D[2k] = Re c[k],  D[M-1-2k] = -Im c[k]
 *   and the TDAC unfolding  y[m] = D[m+M/2] (m<M/2),  -D[3M/2-1-m] (M/2<=m<3M/2),  -D[m-3M/2] (m>=3M/2). */
typedef struct { uint32_t n; float* wre; float* wim; float* fre; float* fim; uint32_t* rev; } fast_plan;
static fast_plan g_plans[16];

static fast_plan* get_plan(uint32_t n) {
	int slot = 0;
	for(uint32_t t = n; t > 1; t >>= 1) ++slot;
	fast_plan* p = &g_plans[slot & 15];
	if(p->n == n) return p;
	uint32_t M = n / 2, Q = M / 2;
	p->wre = (float*) realloc(p->wre, Q * sizeof(float)); p->wim = (float*) realloc(p->wim, Q * sizeof(float));
	p->fre = (float*) realloc(p->fre, Q * sizeof(float)); p->fim = (float*) realloc(p->fim, Q * sizeof(float));
	p->rev = (uint32_t*) realloc(p->rev, Q * sizeof(uint32_t));
	for(uint32_t j = 0; j < Q; ++j) {
		double a = -M_PI * (8.0 * j + 1.0) / (8.0 * M);
		p->wre[j] = (float) cos(a); p->wim[j] = (float) sin(a);
		double b = -2.0 * M_PI * j / Q;
		p->fre[j] = (float) cos(b); p->fim[j] = (float) sin(b);
	}
	int bits = 0;
	for(uint32_t t = Q; t > 1; t >>= 1) ++bits;
	for(uint32_t j = 0; j < Q; ++j) {
		uint32_t r = 0;
		for(int b = 0; b < bits; ++b) if(j & (1u << b)) r |= 1u << (bits - 1 - b);
		p->rev[j] = r;
	}
	p->n = n;
	return p;
}

void por_imdct_fast(const float* in, uint32_t n, float* out) {
	fast_plan* p = get_plan(n);
	const uint32_t M = n / 2, Q = M / 2;
	float re[2048], im[2048]; /* n <= 8192 -> Q <= 2048 */
	for(uint32_t j = 0; j < Q; ++j) {
		float a = in[2 * j], b = in[M - 1 - 2 * j];
		uint32_t r = p->rev[j];
		re[r] = a * p->wre[j] - b * p->wim[j];
		im[r] = a * p->wim[j] + b * p->wre[j];
	}
	for(uint32_t len = 2; len <= Q; len <<= 1) {
		uint32_t half = len / 2, stride = Q / len;
		for(uint32_t s = 0; s < Q; s += len)
			for(uint32_t k = 0; k < half; ++k) {
				float wr = p->fre[k * stride], wi = p->fim[k * stride];
				float xr = re[s + k + half], xi = im[s + k + half];
				float tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
				re[s + k + half] = re[s + k] - tr; im[s + k + half] = im[s + k] - ti;
				re[s + k] += tr; im[s + k] += ti;
			}
	}
	float D[4096];
	for(uint32_t k = 0; k < Q; ++k) {
		float cr = re[k] * p->wre[k] - im[k] * p->wim[k];
		float ci = re[k] * p->wim[k] + im[k] * p->wre[k];
		D[2 * k] = cr; D[M - 1 - 2 * k] = -ci;
	}
	for(uint32_t m = 0; m < M / 2; ++m) out[m] = D[m + M / 2];
	for(uint32_t m = M / 2; m < 3 * M / 2; ++m) out[m] = -D[3 * M / 2 - 1 - m];
	for(uint32_t m = 3 * M / 2; m < n; ++m) out[m] = -D[m - 3 * M / 2];
}

/* ------------------------------------------------------------------------------------------------------------
 * window — src/ParseOggVorbis.hpp:837-862. The reference mixes float and double: x is rounded to float, the
 * products are formed in double, sinf() then takes the float conversion of that double.
 * ---------------------------------------------------------------------------------------------------------- */
/* The reference's own transform as an IMDCT callback without any interpreter in the loop: the binding hands over the
 * addresses of mdct_init / mdct_backward (src/mdct.h:99-105) from oracle/_ref; plans are per thread because the
 * lookup struct (src/mdct.h:87-97: two ints, two table pointers, a float) is built lazily here. */
typedef void (*ref_mdct_init_fn)(void* lookup, int n);
typedef void (*ref_mdct_backward_fn)(void* lookup, float* in, float* out);
static ref_mdct_init_fn g_ref_mdct_init;
static ref_mdct_backward_fn g_ref_mdct_backward;

void por_bind_reference_mdct(void* init_fn, void* backward_fn) {
	g_ref_mdct_init = (ref_mdct_init_fn) init_fn;
	g_ref_mdct_backward = (ref_mdct_backward_fn) backward_fn;
}

void por_imdct_reference(void* user, uint32_t n, const float* in, float* out) {
	static __thread struct { uint32_t n; void* lookup[8]; } plans[8];      /* 64 bytes of lookup storage, pointer aligned */
	static __thread int n_plans;
	(void) user;
	int k = 0;
	while(k < n_plans && plans[k].n != n) ++k;
	if(k == n_plans) {
		if(n_plans == 8) k = 0; else ++n_plans;      /* more than 8 block sizes per thread: recycle (tables leak, test code) */
		memset(&plans[k], 0, sizeof plans[k]);
		plans[k].n = n;
		g_ref_mdct_init(plans[k].lookup, (int) n);
	}
	g_ref_mdct_backward(plans[k].lookup, (float*) in, out);
}

void por_window(uint32_t bs0, uint32_t bs1, int blockflag, int prev, int next, float* out) {
	uint32_t n = blockflag ? bs1 : bs0;
	if(!blockflag) { prev = 0; next = 0; } /* hpp:876: flags only matter for long blocks */
	uint32_t left = (prev ? bs1 : bs0) / 2, right = (next ? bs1 : bs0) / 2;
	uint32_t left_begin = n / 4 - left / 2, right_begin = n - n / 4 - right / 2;
	memset(out, 0, n * sizeof(float));
	for(uint32_t i = 0; i < left; ++i) {
		float x = sinf((float) (M_PI_2 * ((int) i + 0.5) / left));
		out[left_begin + i] = sinf((float) (M_PI_2 * x * x));
	}
	for(uint32_t i = left_begin + left; i < right_begin; ++i) out[i] = 1.0f;
	for(uint32_t i = 0; i < right; ++i) {
		float x = sinf((float) (M_PI_2 * ((int) right - (int) i - .5) / right));
		out[right_begin + i] = sinf((float) (M_PI_2 * x * x));
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * whole batch — src/ParseOggVorbis.hpp:1128-1274 minus the bit reader, plus the overlap-add in the gather form
 * of SURVEY.md §8 a9 (probed bit-exact against VorbisStreamDecodeState, hpp:1008-1109):
 *     PCM[a] = (0 + prev[a-start_prev]*wprev[..]) + cur[a-start_cur]*wcur[..]   (terms outside a frame dropped)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
	float* win[5]; /* [0]=short, [1+prev+2*next]=long variants */
} setup_cache;

static uint32_t floor_posts(const pov_setup* su, const pov_mapping* mp, uint32_t ch) {
	return su->floors[mp->submap_floor[mp->mux[ch]]].n_posts;
}

int por_synth_batch_ex(const pov_setup* setups, uint32_t n_setups, const pov_batch* batch, int imdct,
                       por_imdct_fn ext_imdct, void* user, float* pcm, uint32_t* status,
                       float* cap_after_residue, float* cap_after_envelope, float* cap_pcm_after_mdct,
                       uint32_t cap_stride /* floats per (packet,channel) slot, >= max n */) {
	setup_cache* cache = (setup_cache*) calloc(n_setups, sizeof(setup_cache));
	int rc = POV_OK;
	const uint32_t NMAX = 8192;
	float* spec = (float*) malloc(sizeof(float) * POV_MAX_CHANNELS * (NMAX / 2));
	float* flo = (float*) malloc(sizeof(float) * POV_MAX_CHANNELS * NMAX);
	float* frame = (float*) malloc(sizeof(float) * POV_MAX_CHANNELS * NMAX);
	float* prevw = (float*) malloc(sizeof(float) * POV_MAX_CHANNELS * NMAX); /* previous frame, already windowed */
	float* tmp = (float*) malloc(sizeof(float) * POV_MAX_CHANNELS * (NMAX / 2));
	uint32_t* fl_u = (uint32_t*) malloc(sizeof(uint32_t) * NMAX);
	if(pcm) memset(pcm, 0, sizeof(float) * batch->pcm_floats);

	for(uint32_t si = 0; si < batch->n_streams && rc == POV_OK; ++si) {
		const pov_stream* st = &batch->streams[si];
		if(st->setup_id >= n_setups) { rc = POV_ERR_ARG; break; }
		const pov_setup* su = &setups[st->setup_id];
		setup_cache* sc = &cache[st->setup_id];
		const uint32_t C = su->channels;
		if(!sc->win[0]) {
			sc->win[0] = (float*) malloc(sizeof(float) * su->blocksize[0]);
			por_window(su->blocksize[0], su->blocksize[1], 0, 0, 0, sc->win[0]);
			for(int w = 0; w < 4; ++w) {
				sc->win[1 + w] = (float*) malloc(sizeof(float) * su->blocksize[1]);
				por_window(su->blocksize[0], su->blocksize[1], 1, w & 1, (w >> 1) & 1, sc->win[1 + w]);
			}
		}
		uint32_t n_prev = 0;
		for(uint32_t pi = 0; pi < st->n_packets; ++pi) {
			const uint32_t gp = st->first_packet + pi;
			const pov_packet* pk = &batch->packets[gp];
			if(pk->mode >= su->n_modes) { rc = POV_ERR_ARG; break; }
			const pov_mode* md = &su->modes[pk->mode];
			const pov_mapping* mp = &su->mappings[md->mapping];
			const uint32_t n = su->blocksize[md->blockflag ? 1 : 0], half = n / 2;
			const float* win = md->blockflag ? sc->win[1 + (pk->window_flags & 3)] : sc->win[0];
			uint32_t pst = 0;
			uint8_t used[POV_MAX_CHANNELS];

			/* 4.3.2 floor curves (hpp:1158-1172) */
			uint64_t yo = pk->ys_off;
			for(uint32_t c = 0; c < C; ++c) {
				used[c] = (pk->floor_used >> c) & 1;
				float* f = flo + (size_t) c * NMAX;
				memset(f, 0, sizeof(float) * n); /* hpp:1159: floor_outputs zero-initialised */
				if(!used[c]) continue;
				const pov_floor1* fl = &su->floors[mp->submap_floor[mp->mux[c]]];
				uint32_t ys32[POV_MAX_POSTS], fy[POV_MAX_POSTS];
				uint8_t flag[POV_MAX_POSTS];
				for(uint32_t i = 0; i < fl->n_posts; ++i) ys32[i] = batch->ys[yo + i];
				yo += fl->n_posts;
				pst |= por_floor1_unwrap(fl->xs, fl->n_posts, fl->multiplier, ys32, fy, flag);
				por_floor1_render(fl->xs, fl->n_posts, fl->multiplier, fy, flag, n, fl_u);
				pst |= por_floor1_db(fl_u, n, f);
			}
			(void) floor_posts;
			/* 4.3.3 nonzero propagate (hpp:1174-1180) */
			for(uint32_t k = 0; k < mp->n_couplings; ++k)
				if(used[mp->coupling_ang[k]] || used[mp->coupling_mag[k]])
					used[mp->coupling_ang[k]] = used[mp->coupling_mag[k]] = 1;

			/* 4.3.4 residue (hpp:1182-1211) */
			if(batch->input_kind == POV_INPUT_DENSE) {
				memcpy(spec, (const float*) batch->payload + pk->spec_off, sizeof(float) * C * half);
			} else {
				const uint8_t* pl = (const uint8_t*) batch->payload + pk->spec_off;
				int ebits = 16;
				for(uint32_t b = 0; b < su->n_codebooks; ++b) if(su->codebooks[b].n_entries > 65536) ebits = 32;
				memset(spec, 0, sizeof(float) * C * half);
				for(uint32_t s = 0; s < mp->n_submaps; ++s) {
					uint32_t chs[POV_MAX_CHANNELS], nch = 0;
					uint8_t su_used[POV_MAX_CHANNELS];
					for(uint32_t c = 0; c < C; ++c) if(mp->mux[c] == s) { su_used[nch] = used[c]; chs[nch++] = c; }
					const pov_residue* rs = &su->residues[mp->submap_residue[s]];
					uint64_t used_bytes = 0;
					memset(tmp, 0, sizeof(float) * nch * half);
					if(rs->type == 2) { /* hpp:685-694: one interleaved vector, always "used" */
						uint8_t one = 1;
						pst |= por_residue_apply(rs, su->codebooks, su->n_codebooks, 1, &one, nch * half, ebits, pl,
						                         &used_bytes, tmp);
						for(uint32_t j = 0; j < nch; ++j)
							for(uint32_t i = 0; i < half; ++i) spec[(size_t) chs[j] * half + i] = tmp[j + nch * i];
					} else {
						pst |= por_residue_apply(rs, su->codebooks, su->n_codebooks, nch, su_used, half, ebits, pl,
						                         &used_bytes, tmp);
						for(uint32_t j = 0; j < nch; ++j)
							memcpy(spec + (size_t) chs[j] * half, tmp + (size_t) j * half, sizeof(float) * half);
					}
					pl += used_bytes;
				}
			}
			if(cap_after_residue)
				for(uint32_t c = 0; c < C; ++c)
					memcpy(cap_after_residue + ((size_t) gp * C + c) * cap_stride, spec + (size_t) c * half, sizeof(float) * half);

			/* 4.3.5 inverse coupling, last step first (hpp:1213-1241) */
			for(uint32_t k = mp->n_couplings; k > 0; --k)
				por_inverse_coupling(spec + (size_t) mp->coupling_mag[k - 1] * half,
				                     spec + (size_t) mp->coupling_ang[k - 1] * half, half);
			/* 4.3.6 dot product (hpp:1243-1255) */
			for(uint32_t c = 0; c < C; ++c) {
				if(used[c]) {
					float* r = spec + (size_t) c * half;
					const float* f = flo + (size_t) c * NMAX;
					for(uint32_t i = 0; i < half; ++i) r[i] *= f[i];
				}
				if(cap_after_envelope)
					memcpy(cap_after_envelope + ((size_t) gp * C + c) * cap_stride, spec + (size_t) c * half, sizeof(float) * half);
			}
			/* 4.3.7 inverse MDCT (hpp:1257-1265) */
			for(uint32_t c = 0; c < C; ++c) {
				float* y = frame + (size_t) c * NMAX;
				const float* X = spec + (size_t) c * half;
				if(imdct == 0) por_imdct_closed(X, n, y);
				else if(imdct == 1) por_imdct_fast(X, n, y);
				else ext_imdct(user, n, X, y);
				if(cap_pcm_after_mdct)
					memcpy(cap_pcm_after_mdct + ((size_t) gp * C + c) * cap_stride, y, sizeof(float) * n);
			}
			if(status) status[gp] = pst;

			/* window + overlap-add + emit (hpp:1008-1059 in gather form) */
			if(pi > 0 && pk->emit_frames > 0 && pcm) {
				const int64_t shift = (int64_t) n / 4 - (int64_t) n_prev / 4; /* centre(prev) - start(cur) */
				for(uint32_t c = 0; c < C; ++c) {
					const float* pw = prevw + (size_t) c * NMAX;
					const float* y = frame + (size_t) c * NMAX;
					for(uint32_t j = 0; j < pk->emit_frames; ++j) {
						float acc = 0.0f;
						uint32_t ip = n_prev / 2 + j;
						if(ip < n_prev) acc = acc + pw[ip];
						int64_t ic = (int64_t) j + shift;
						if(ic >= 0 && ic < (int64_t) n) acc = acc + y[ic] * win[ic];
						uint64_t f = pk->pcm_off + j;
						uint64_t o = (batch->pcm_layout == POV_PCM_PLANAR)
							? st->pcm_base + (uint64_t) c * st->pcm_frames + f
							: st->pcm_base + f * C + c;
						pcm[o] = acc;
					}
				}
			}
			for(uint32_t c = 0; c < C; ++c) {
				float* pw = prevw + (size_t) c * NMAX;
				const float* y = frame + (size_t) c * NMAX;
				for(uint32_t i = 0; i < n; ++i) pw[i] = 0.0f + y[i] * win[i];
			}
			n_prev = n;
		}
	}
	for(uint32_t i = 0; i < n_setups; ++i) for(int w = 0; w < 5; ++w) free(cache[i].win[w]);
	free(cache); free(spec); free(flo); free(frame); free(prevw); free(tmp); free(fl_u);
	return rc;
}

int por_synth_batch(const pov_setup* setups, uint32_t n_setups, const pov_batch* batch, int imdct,
                    por_imdct_fn ext_imdct, void* user, float* pcm, uint32_t* status) {
	return por_synth_batch_ex(setups, n_setups, batch, imdct, ext_imdct, user, pcm, status, NULL, NULL, NULL, 0);
}
