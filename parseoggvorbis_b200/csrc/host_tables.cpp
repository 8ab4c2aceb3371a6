// Host-side table generation for libpov_synth.so. Every table is derived from the Vorbis I definitions the
// reference implements; citations are relative to the reference root.
#include "host_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace pov {

// src/inverse_db_table.h:13-78 holds the 256 literals of Vorbis I spec 10.1. The spec table is
// fromdB((i-255)*0.546875), fromdB(x) = exp(x*0.11512925), printed with 8 significant digits; the float nearest
// to that 8-digit decimal is what every decoder compiles in. Generating it the same way is bit-identical to the
// reference's literals (pinned by tests/test_abi.py::test_inverse_db_table_matches_reference_golden).
void make_inverse_db_table(float out[256]) {
	for(int i = 0; i < 256; ++i) {
		char buf[64];
		const double v = exp((double) (i - 255) * 0.546875 * 0.11512925);
		snprintf(buf, sizeof buf, "%.7e", v);
		out[i] = strtof(buf, nullptr);
	}
}

// src/ParseOggVorbis.hpp:850-853: x is rounded to float, products are formed in double, sinf takes the float
// conversion of the double argument. The falling slope (hpp:856-859) evaluates the same expression at the
// mirrored index, so only the rising slope is stored.
void make_window_slope(uint32_t len, std::vector<float>& out) {
	out.resize(len);
	for(uint32_t i = 0; i < len; ++i) {
		const float x = sinf((float) (M_PI_2 * ((int) i + 0.5) / (int) len));
		out[i] = sinf((float) (M_PI_2 * x * x));
	}
}

void make_window(uint32_t bs0, uint32_t bs1, int blockflag, int prev, int next, std::vector<float>& out) {
	const uint32_t n = blockflag ? bs1 : bs0;
	if(!blockflag) prev = next = 0;                    // hpp:876: flags matter for long blocks only
	const uint32_t left = (prev ? bs1 : bs0) / 2, right = (next ? bs1 : bs0) / 2;
	const uint32_t lb = n / 4 - left / 2, rb = n - n / 4 - right / 2;
	std::vector<float> sl, sr;
	make_window_slope(left, sl);
	make_window_slope(right, sr);
	out.assign(n, 0.f);
	for(uint32_t i = 0; i < left; ++i) out[lb + i] = sl[i];
	for(uint32_t i = lb + left; i < rb; ++i) out[i] = 1.f;
	for(uint32_t i = 0; i < right; ++i) out[rb + i] = sr[right - 1 - i];
}

void make_rotation(uint32_t n, std::vector<float>& out) {
	const uint32_t M = n / 2, Q = n / 4;
	out.resize(2 * (size_t) Q);
	for(uint32_t j = 0; j < Q; ++j) {
		const double a = -M_PI * (8.0 * j + 1.0) / (8.0 * M);
		out[2 * j] = (float) cos(a);
		out[2 * j + 1] = (float) sin(a);
	}
}

void make_fft_twiddles(uint32_t n, std::vector<float>& out) {
	const uint32_t Q = n / 4;
	out.resize(2 * (size_t) Q);
	for(uint32_t e = 0; e < Q; ++e) {
		const double a = -2.0 * M_PI * e / Q;
		out[2 * e] = (float) cos(a);
		out[2 * e + 1] = (float) sin(a);
	}
}

// Layout must match PassTables<Q> in fft_core.cuh: [first small-radix pass][radix-8 pass L1][radix-8 pass L1/8]...
// radix-4 first pass: Q/4 butterflies x 4 factors (W^0, W^j, W^2j, W^3j); radix-2 first pass: Q/2 x 2 (W^0, W^j);
// radix-8 pass of length L (>= 64): L/8 butterflies x 8 factors W_L^(j*k).
void make_fft_pass_tables(uint32_t n, std::vector<float>& out) {
	const uint32_t Q = n / 4;
	uint32_t log2q = 0;
	while((1u << log2q) < Q) ++log2q;
	const uint32_t r0 = (log2q % 3 == 0) ? 8 : (log2q % 3 == 1) ? 2 : 4;
	out.clear();
	auto put = [&](double num, double den) {
		const double a = -2.0 * M_PI * num / den;
		out.push_back((float) cos(a));
		out.push_back((float) sin(a));
	};
	uint32_t L = Q;
	if(r0 != 8) {
		for(uint32_t j = 0; j < Q / r0; ++j)
			for(uint32_t k = 0; k < r0; ++k) put((double) j * k, (double) Q);
		L = Q / r0;
	}
	for(; L >= 64; L /= 8)
		for(uint32_t j = 0; j < L / 8; ++j)
			for(uint32_t k = 0; k < 8; ++k) put((double) j * k, (double) L);
	if(out.empty()) { out.push_back(1.f); out.push_back(0.f); }
}

// Layout must match Tw8Tables<Q> in fft_core.cuh: radix-8 passes in execution order (L = L1, L1/8, ... >= 64), per
// pass the factors W_L^(j*1..4) of butterfly j, split in two halves (see below).
void make_fft_r8_tables(uint32_t n, std::vector<float>& out) {
	const uint32_t Q = n / 4;
	uint32_t log2q = 0;
	while((1u << log2q) < Q) ++log2q;
	const uint32_t r0 = (log2q % 3 == 0) ? 8 : (log2q % 3 == 1) ? 2 : 4;
	out.clear();
	// per pass: [L/8 x (W^j, W^2j)] then [L/8 x (W^3j, W^4j)] — a warp whose lanes take consecutive butterflies reads each
	// half with consecutive 16-byte loads (no shared-memory bank conflicts)
	for(uint32_t L = (r0 == 8) ? Q : Q / r0; L >= 64; L /= 8)
		for(uint32_t half = 0; half < 2; ++half)
			for(uint32_t j = 0; j < L / 8; ++j)
				for(uint32_t k = 1 + 2 * half; k <= 2 + 2 * half; ++k) {
					const double a = -2.0 * M_PI * (double) (j * k) / (double) L;
					out.push_back((float) cos(a));
					out.push_back((float) sin(a));
				}
	if(out.empty()) { out.assign(8, 0.f); }
}

void make_rotation_consts(uint32_t n, float c1[2], float c6[2]) {
	const double M = n / 2;
	c1[0] = (float) cos(-M_PI / M); c1[1] = (float) sin(-M_PI / M);
	c6[0] = (float) cos(M_PI * 6.0 / (8.0 * M)); c6[1] = (float) sin(M_PI * 6.0 / (8.0 * M));
}

// ---- tensor-memory lane tables of the 512-point FFT -------------------------------------------------------------------
// The FFT (kernel_warp.cu fft512_tm; executable model with the measured tcgen05 shape maps: tools/model/tmem_fft_model.py):
//   pass 1  radix 8 over j8 j7 j6, lane = (j5 j4 j3 j2 j1), registers 2 k0 + j0
//   pass 2  radix 4 over j5 j4,    lane = (j3 j2 j1 | k0_1^r k0_0^r), r = k0_2
//   pass 3  radix 4 over j3 j2,    lane = (j1 . . | k1_1^r k1_0^r)
//   pass 4  radix 4 over j1 j0,    lane t = (k0_1^r k1_1^r k1_0^r k2_1^r k2_0^r), register R = 8 k3_1 + 4 r + 2 k3_0 + (k0_0 ^ r)
// output frequency k = k0 + 8 k1 + 32 k2 + 128 k3.
uint32_t tm_fft_freq_of(uint32_t lane, uint32_t reg) {
	const uint32_t k3 = ((reg >> 3) << 1) | ((reg >> 1) & 1), r = (reg >> 2) & 1, h = reg & 1, x = r ? 1u : 0u;
	const uint32_t k0_0 = h ^ x, k0_1 = ((lane >> 4) & 1) ^ x, k0_2 = r;
	const uint32_t k1 = ((lane >> 2) & 3) ^ (x * 3), k2 = (lane & 3) ^ (x * 3);
	return k0_0 | (k0_1 << 1) | (k0_2 << 2) | (k1 << 3) | (k2 << 5) | (k3 << 7);
}

void make_tm_lane_tables(uint32_t n, const std::vector<float>& rot, const std::vector<float>& slope, std::vector<float>& out) {
	const uint32_t Q = n / 4;                 // 512
	out.assign(32 * (size_t) kTmTableCols, 0.f);
	auto W = [](double e, double N, float* dst) { const double a = -2.0 * M_PI * e / N; dst[0] = (float) cos(a); dst[1] = (float) sin(a); };
	for(uint32_t l = 0; l < 32; ++l) {
		float* row = &out[(size_t) l * kTmTableCols];
		// [0,32): spectral stage, quads q = l + 32 m and Q/2 - 1 - q: rotation pairs (w[2q], w[2q+1])
		for(uint32_t m = 0; m < 4; ++m) {
			const uint32_t q = l + 32 * m, q2 = Q / 2 - 1 - q;
			for(uint32_t i = 0; i < 4; ++i) { row[8 * m + i] = rot[4 * q + i]; row[8 * m + 4 + i] = rot[4 * q2 + i]; }
		}
		// [32,48): pass 1 twiddles W_512^(e c), c = 1..4, e = 2 l + j0
		for(uint32_t j0 = 0; j0 < 2; ++j0)
			for(uint32_t c = 1; c <= 4; ++c) W((double) ((2 * l + j0) * c), 512.0, row + 32 + 8 * j0 + 2 * (c - 1));
		// [48,64): pass 2 twiddles W_64^(e c), c = 1..3, e = (j3 j2 j1 j0) = 2 (l >> 2) + j0
		for(uint32_t j0 = 0; j0 < 2; ++j0)
			for(uint32_t c = 1; c <= 3; ++c) W((double) ((2 * (l >> 2) + j0) * c), 64.0, row + 48 + 8 * j0 + 2 * (c - 1));
		// [64,80): pass 3 twiddles W_16^(e c), c = 1..3, e = (j1 j0) = 2 (l >> 4) + j0
		for(uint32_t j0 = 0; j0 < 2; ++j0)
			for(uint32_t c = 1; c <= 3; ++c) W((double) ((2 * (l >> 4) + j0) * c), 16.0, row + 64 + 8 * j0 + 2 * (c - 1));
		// [80,112): post-rotation w[k] of the 16 outputs in register order
		for(uint32_t R = 0; R < 16; ++R) {
			const uint32_t k = tm_fft_freq_of(l, R);
			row[80 + 2 * R] = rot[2 * k]; row[80 + 2 * R + 1] = rot[2 * k + 1];
		}
		// [112,144): overlap-add, iteration i: rising slope at j..j+3 and at 2Q-4-j..2Q-1-j, j = 4 * (true quad of storage quad 32 i + l)
		for(uint32_t i = 0; i < 4; ++i) {
			const uint32_t sq = 32 * i + l;
			const uint32_t tq = (sq & 0x44u) | ((sq & 3u) << 4) | ((sq & 0x20u) >> 2) | ((sq & 0x18u) >> 3);
			const uint32_t j = 4 * tq;
			for(uint32_t e = 0; e < 4; ++e) { row[112 + 8 * i + e] = slope[j + e]; row[112 + 8 * i + 4 + e] = slope[2 * Q - 4 - j + e]; }
		}
	}
}

// Neighbours (src/Utils.hpp:60-118) depend on the X list only, so they are found once here instead of once per
// packet and post as in the reference (hpp:532-533). level[] orders the posts so that a post's two neighbours
// are final before it is unwrapped; sorted_idx is the ascending-x order of hpp:458-469.
bool make_floor_tables(const pov_floor1& in, DevFloor& out, std::string& msg) {
	memset(&out, 0, sizeof out);
	const int posts = in.n_posts;
	if(posts < 2 || posts > (int) POV_MAX_POSTS) { msg = "floor1: n_posts out of range"; return false; }
	if(in.multiplier < 1 || in.multiplier > 4) { msg = "floor1: multiplier out of range (hpp:491)"; return false; }
	if(in.xs[0] != 0) { msg = "floor1: xs[0] must be 0 (hpp:449)"; return false; }
	for(int i = 2; i < posts; ++i)
		if(in.xs[i] == 0 || in.xs[i] >= in.xs[1]) { msg = "floor1: xs[i] outside (0, xs[1]) (hpp:450,455)"; return false; }
	{
		std::vector<uint16_t> s(in.xs, in.xs + posts);
		std::sort(s.begin(), s.end());
		if(std::adjacent_find(s.begin(), s.end()) != s.end()) { msg = "floor1: duplicate X values (Utils.hpp:145 asserts x0 < x1)"; return false; }
	}
	static const uint32_t ranges[4] = {256, 128, 86, 64};
	out.n_posts = (uint16_t) posts;
	out.multiplier = in.multiplier;
	out.range = ranges[in.multiplier - 1];
	int maxlevel = 0;
	for(int i = 0; i < posts; ++i) {
		out.xs[i] = in.xs[i];
		if(i < 2) { out.level[i] = 0; continue; }
		int lo = -1, hi = -1;
		for(int j = 0; j < i; ++j) {
			if(in.xs[j] < in.xs[i] && (lo < 0 || in.xs[j] > in.xs[lo])) lo = j;
			if(in.xs[j] > in.xs[i] && (hi < 0 || in.xs[j] < in.xs[hi])) hi = j;
		}
		out.lo[i] = (uint8_t) lo;
		out.hi[i] = (uint8_t) hi;
		out.dxn[i] = (uint16_t) (in.xs[i] - in.xs[lo]);
		out.adx[i] = (uint16_t) (in.xs[hi] - in.xs[lo]);
		out.rinv[i] = 1.0f / (float) out.adx[i];
		const int lv = 1 + std::max<int>(out.level[lo], out.level[hi]);
		out.level[i] = (uint8_t) lv;
		maxlevel = std::max(maxlevel, lv);
	}
	out.n_levels = (uint8_t) (maxlevel + 1);
	std::vector<int> order(posts);
	for(int i = 0; i < posts; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return in.xs[a] < in.xs[b]; });
	for(int i = 0; i < posts; ++i) out.sorted_idx[i] = (uint8_t) order[i];
	return true;
}

}  // namespace pov
