// Inverse-MDCT building blocks (sm_100a): packed-f32x2 complex arithmetic, radix-8/4/2 butterflies and the
// shared-memory pass driver of the Q = n/4 point complex FFT that evaluates a DCT-IV.
//
// Algorithm (ours; the reference's src/mdct.cpp — libvorbis' split-radix code — is deliberately NOT followed):
//   contract (src/mdct.cpp:433-527, probed):  y[m] = sum_k X[k] cos((2pi/n)(m + 1/2 + n/4)(k + 1/2)), m < n
//   M = n/2, Q = n/4.   t[j] = (X[2j] + i X[M-1-2j]) * w[j],   w[j] = exp(-i pi (8j+1) / (8M))
//                       T = FFT_Q(t);   c[k] = T[k] * w[k];     D[2k] = Re c[k],  D[M-1-2k] = -Im c[k]
//   y[m] = D[m+M/2] (m < M/2),  -D[3M/2-1-m] (M/2 <= m < 3M/2),  -D[m-3M/2] (m >= 3M/2)
// so a frame is fully described by the M values D[], its first half by D[M/2..M), its second by D[0..M/2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pov {

// ---- complex helpers on float2 (x = re, y = im); FADD2/FMUL2/FFMA2 on sm_100 ------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
	const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(w.y, w.x));  // ai*wi, ai*wr
	return __ffma2_rn(make_float2(a.x, a.x), w, make_float2(-t.x, t.y));        // ar*wr - ai*wi, ar*wi + ai*wr
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }    // a * (-i)
#define POV_SQRT1_2 0.70710678118654752440f
__device__ __forceinline__ float2 mul_w8_1(float2 a) {  // a * (1 - i)/sqrt2
	return __fmul2_rn(make_float2(a.x + a.y, a.y - a.x), make_float2(POV_SQRT1_2, POV_SQRT1_2));
}
__device__ __forceinline__ float2 mul_w8_3(float2 a) {  // a * (-1 - i)/sqrt2
	return __fmul2_rn(make_float2(a.y - a.x, -(a.x + a.y)), make_float2(POV_SQRT1_2, POV_SQRT1_2));
}

// forward DFTs, outputs in natural order (a[k] = X[k])
__device__ __forceinline__ void dft2(float2& a0, float2& a1) {
	float2 s = cadd(a0, a1), d = csub(a0, a1);
	a0 = s; a1 = d;
}
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
	float2 p0 = cadd(a0, a2), p1 = csub(a0, a2), p2 = cadd(a1, a3), p3 = mul_mi(csub(a1, a3));
	a0 = cadd(p0, p2); a2 = csub(p0, p2); a1 = cadd(p1, p3); a3 = csub(p1, p3);
}
__device__ __forceinline__ void dft8(float2* a) {
	float2 u0 = cadd(a[0], a[4]), v0 = csub(a[0], a[4]);
	float2 u1 = cadd(a[1], a[5]), v1 = mul_w8_1(csub(a[1], a[5]));
	float2 u2 = cadd(a[2], a[6]), v2 = mul_mi(csub(a[2], a[6]));
	float2 u3 = cadd(a[3], a[7]), v3 = mul_w8_3(csub(a[3], a[7]));
	dft4(u0, u1, u2, u3);   // -> X0, X2, X4, X6
	dft4(v0, v1, v2, v3);   // -> X1, X3, X5, X7
	a[0] = u0; a[2] = u1; a[4] = u2; a[6] = u3;
	a[1] = v0; a[3] = v1; a[5] = v2; a[7] = v3;
}

// ---- shared-memory layout of one FFT work buffer -------------------------------------------------------------
// Element i of a Q-point buffer lives at pad(i) = i + i/8 (float2 units): one pad slot per 8 elements makes the
// stride-8^k accesses of every radix-8 pass fall on distinct banks (analysis in DESIGN.md §IMDCT).
__device__ __forceinline__ int tpad(int i) { return i + (i >> 3); }
template <int Q> struct FftGeom {
	static constexpr int kStride = Q + Q / 8;          // float2 slots per FFT buffer
	static constexpr int kItems = Q / 8;               // 8-point work items per FFT per pass
	static constexpr int kLog2 = (Q == 16) ? 4 : (Q == 32) ? 5 : (Q == 64) ? 6 : (Q == 128) ? 7 : (Q == 256) ? 8
	                           : (Q == 512) ? 9 : (Q == 1024) ? 10 : 11;
	static constexpr int kFirstRadix = (kLog2 % 3 == 0) ? 8 : (kLog2 % 3 == 1) ? 2 : 4;   // [r, 8, 8, ...]
	static constexpr int kRadix8Passes = kLog2 / 3;
};

// One radix-8 DIF pass over sub-FFTs of length L inside a Q-point buffer (in place: the item reads and writes
// the same 8 slots). W = table of W_Q^e, e < Q.
template <int Q, int L>
__device__ __forceinline__ void pass_radix8(float2* __restrict__ T, int t, const float2* __restrict__ W) {
	constexpr int s = L / 8;
	const int blk = t / s, j = t - blk * s;
	const int base = blk * L + j;
	float2 a[8];
#pragma unroll
	for(int m = 0; m < 8; ++m) a[m] = T[tpad(base + m * s)];
	dft8(a);
	if(L > 8) {
		constexpr int step = Q / L;
#pragma unroll
		for(int k = 1; k < 8; ++k) a[k] = cmul(a[k], __ldg(&W[(step * j * k) & (Q - 1)]));
	}
#pragma unroll
	for(int k = 0; k < 8; ++k) T[tpad(base + k * s)] = a[k];
}

// First pass when log2(Q) is not a multiple of 3: radix 4 (two butterflies per item) or radix 2 (four).
template <int Q>
__device__ __forceinline__ void pass_first_small(float2* __restrict__ T, int t, const float2* __restrict__ W) {
	constexpr int R = FftGeom<Q>::kFirstRadix;
	if(R == 4) {
		constexpr int s = Q / 4;
#pragma unroll
		for(int u = 0; u < 2; ++u) {
			const int j = t + u * (Q / 8);
			float2 a0 = T[tpad(j)], a1 = T[tpad(j + s)], a2 = T[tpad(j + 2 * s)], a3 = T[tpad(j + 3 * s)];
			dft4(a0, a1, a2, a3);
			a1 = cmul(a1, __ldg(&W[j]));
			a2 = cmul(a2, __ldg(&W[(2 * j) & (Q - 1)]));
			a3 = cmul(a3, __ldg(&W[(3 * j) & (Q - 1)]));
			T[tpad(j)] = a0; T[tpad(j + s)] = a1; T[tpad(j + 2 * s)] = a2; T[tpad(j + 3 * s)] = a3;
		}
	} else if(R == 2) {
		constexpr int s = Q / 2;
#pragma unroll
		for(int u = 0; u < 4; ++u) {
			const int j = t + u * (Q / 8);
			float2 a0 = T[tpad(j)], a1 = T[tpad(j + s)];
			dft2(a0, a1);
			a1 = cmul(a1, __ldg(&W[j]));
			T[tpad(j)] = a0; T[tpad(j + s)] = a1;
		}
	}
}

// Frequency index held at buffer position p after all DIF passes (mixed-radix digit reversal).
template <int Q>
__device__ __forceinline__ int freq_of_pos(int p) {
	constexpr int R0 = FftGeom<Q>::kFirstRadix;
	int k = 0, mul = 1, rem = p, L = Q;
	if(R0 != 8) {
		const int s = Q / R0, d = rem / s;
		rem -= d * s; k += d; mul = R0; L = s;
	}
#pragma unroll
	for(int i = 0; i < FftGeom<Q>::kRadix8Passes; ++i) {
		const int s = L / 8, d = rem / s;
		rem -= d * s; k += d * mul; mul *= 8; L = s;
	}
	return k;
}

// All passes but the last for `nf` FFTs laid out back to back (stride FftGeom<Q>::kStride) — block-wide,
// contains __syncthreads(). After it returns, the last radix-8 pass (L = 8) remains to be done by the caller
// (pass_last_*), which fuses the post-rotation.
template <int Q, int L>
__device__ __forceinline__ void passes_radix8_down_to_16(float2* T, int nf, const float2* W) {
	if constexpr(L >= 64) {
		for(int w = threadIdx.x; w < nf * FftGeom<Q>::kItems; w += blockDim.x) {
			const int f = w / FftGeom<Q>::kItems, t = w - f * FftGeom<Q>::kItems;
			pass_radix8<Q, L>(T + f * FftGeom<Q>::kStride, t, W);
		}
		__syncthreads();
		passes_radix8_down_to_16<Q, L / 8>(T, nf, W);
	}
}

template <int Q>
__device__ __forceinline__ void fft_passes_except_last(float2* T, int nf, const float2* W) {
	constexpr int R0 = FftGeom<Q>::kFirstRadix;
	if constexpr(R0 != 8) {
		for(int w = threadIdx.x; w < nf * FftGeom<Q>::kItems; w += blockDim.x) {
			const int f = w / FftGeom<Q>::kItems, t = w - f * FftGeom<Q>::kItems;
			pass_first_small<Q>(T + f * FftGeom<Q>::kStride, t, W);
		}
		__syncthreads();
		passes_radix8_down_to_16<Q, Q / R0>(T, nf, W);
	} else {
		passes_radix8_down_to_16<Q, Q>(T, nf, W);
	}
}

// Last pass (L = 8, no FFT twiddles) of item t fused with the DCT-IV post-rotation: writes D[2k] and D[M-1-2k]
// (M = 2Q) of its 8 frequencies into the float array Dst (plain layout).
template <int Q>
__device__ __forceinline__ void pass_last_to_D(const float2* __restrict__ T, int t, const float2* __restrict__ rot,
                                               float* __restrict__ Dst) {
	constexpr int M = 2 * Q;
	float2 a[8];
#pragma unroll
	for(int m = 0; m < 8; ++m) a[m] = T[tpad(8 * t + m)];
	dft8(a);
	const int k0 = freq_of_pos<Q>(8 * t);
#pragma unroll
	for(int m = 0; m < 8; ++m) {
		const int k = k0 + m * (Q / 8);
		const float2 c = cmul(a[m], __ldg(&rot[k]));
		Dst[2 * k] = c.x;
		Dst[M - 1 - 2 * k] = -c.y;
	}
}

// ---- per-pass packed twiddle tables (host: make_fft_pass_tables) ------------------------------------------------
template <int Q> struct PassTables {
	static constexpr int R0 = FftGeom<Q>::kFirstRadix;
	static constexpr int kFirstSize = (R0 == 8) ? 0 : Q;             // (Q/R0) butterflies x R0 factors
	static constexpr int kL1 = (R0 == 8) ? Q : Q / R0;               // length of the first radix-8 pass
	static __host__ __device__ constexpr int offset(int L) {         // float2 offset of the radix-8 pass of length L
		int o = kFirstSize;
		for(int l = kL1; l > L; l /= 8) o += l;                       // each radix-8 table holds (l/8)*8 = l factors
		return o;
	}
};

template <int Q, int L>
__device__ __forceinline__ void pass_radix8_p(float2* __restrict__ T, int t, const float2* __restrict__ TWP) {
	constexpr int s = L / 8, ps = s + s / 8;                         // s % 8 == 0 for L >= 64
	const int blk = t / s, j = t - blk * s;
	const int base = blk * L + j;
	float2* p = T + base + (base >> 3);
	float2 a[8];
#pragma unroll
	for(int m = 0; m < 8; ++m) a[m] = p[m * ps];
	dft8(a);
	constexpr int kOff = PassTables<Q>::offset(L);
	const float4* tw = reinterpret_cast<const float4*>(TWP + kOff + j * 8);
	const float4 w01 = __ldg(tw), w23 = __ldg(tw + 1), w45 = __ldg(tw + 2), w67 = __ldg(tw + 3);
	a[1] = cmul(a[1], make_float2(w01.z, w01.w));
	a[2] = cmul(a[2], make_float2(w23.x, w23.y));
	a[3] = cmul(a[3], make_float2(w23.z, w23.w));
	a[4] = cmul(a[4], make_float2(w45.x, w45.y));
	a[5] = cmul(a[5], make_float2(w45.z, w45.w));
	a[6] = cmul(a[6], make_float2(w67.x, w67.y));
	a[7] = cmul(a[7], make_float2(w67.z, w67.w));
#pragma unroll
	for(int k = 0; k < 8; ++k) p[k * ps] = a[k];
}

template <int Q>
__device__ __forceinline__ void pass_first_small_p(float2* __restrict__ T, int t, const float2* __restrict__ TWP) {
	constexpr int R = FftGeom<Q>::kFirstRadix;
	if(R == 4) {
		constexpr int s = Q / 4, ps = s + s / 8;
#pragma unroll
		for(int u = 0; u < 2; ++u) {
			const int j = t + u * (Q / 8);
			float2* p = T + j + (j >> 3);
			float2 a0 = p[0], a1 = p[ps], a2 = p[2 * ps], a3 = p[3 * ps];
			dft4(a0, a1, a2, a3);
			const float4* tw = reinterpret_cast<const float4*>(TWP + j * 4);
			const float4 w01 = __ldg(tw), w23 = __ldg(tw + 1);
			a1 = cmul(a1, make_float2(w01.z, w01.w));
			a2 = cmul(a2, make_float2(w23.x, w23.y));
			a3 = cmul(a3, make_float2(w23.z, w23.w));
			p[0] = a0; p[ps] = a1; p[2 * ps] = a2; p[3 * ps] = a3;
		}
	} else if(R == 2) {
		constexpr int s = Q / 2, ps = s + s / 8;
#pragma unroll
		for(int u = 0; u < 4; ++u) {
			const int j = t + u * (Q / 8);
			float2* p = T + j + (j >> 3);
			float2 a0 = p[0], a1 = p[ps];
			dft2(a0, a1);
			const float4 w01 = __ldg(reinterpret_cast<const float4*>(TWP + j * 2));
			a1 = cmul(a1, make_float2(w01.z, w01.w));
			p[0] = a0; p[ps] = a1;
		}
	}
}

template <int Q, int L>
__device__ __forceinline__ void passes_radix8_p(float2* T, int nf, const float2* TWP) {
	if constexpr(L >= 64) {
		for(int w = threadIdx.x; w < nf * FftGeom<Q>::kItems; w += blockDim.x) {
			const int f = w / FftGeom<Q>::kItems, t = w - f * FftGeom<Q>::kItems;
			pass_radix8_p<Q, L>(T + f * FftGeom<Q>::kStride, t, TWP);
		}
		__syncthreads();
		passes_radix8_p<Q, L / 8>(T, nf, TWP);
	}
}

// All passes but the last (block-wide, contains __syncthreads()), packed-table version.
template <int Q>
__device__ __forceinline__ void fft_passes_except_last_p(float2* T, int nf, const float2* TWP) {
	if constexpr(FftGeom<Q>::kFirstRadix != 8) {
		for(int w = threadIdx.x; w < nf * FftGeom<Q>::kItems; w += blockDim.x) {
			const int f = w / FftGeom<Q>::kItems, t = w - f * FftGeom<Q>::kItems;
			pass_first_small_p<Q>(T + f * FftGeom<Q>::kStride, t, TWP);
		}
		__syncthreads();
	}
	passes_radix8_p<Q, PassTables<Q>::kL1>(T, nf, TWP);
}

// Last pass + post-rotation, index arithmetic folded: positions 8t..8t+7 live at T[9t + m].
template <int Q>
__device__ __forceinline__ void pass_last_to_D_p(const float2* __restrict__ T, int t, const float2* __restrict__ rot,
                                                 float* __restrict__ Dst) {
	constexpr int M = 2 * Q;
	float2 a[8];
	const float2* p = T + 9 * t;
#pragma unroll
	for(int m = 0; m < 8; ++m) a[m] = p[m];
	dft8(a);
	const int k0 = freq_of_pos<Q>(8 * t);
	const float2* r = rot + k0;
	float* d0 = Dst + 2 * k0;
	float* d1 = Dst + (M - 1 - 2 * k0);
#pragma unroll
	for(int m = 0; m < 8; ++m) {
		const float2 c = cmul(a[m], __ldg(r + m * (Q / 8)));
		d0[m * (Q / 4)] = c.x;
		d1[-m * (Q / 4)] = -c.y;
	}
}

// ---- compact radix-8 tables held in SHARED memory by the fused kernel (host: make_fft_r8_tables) -----------------
template <int Q> struct Tw8Tables {
	static constexpr int kL1 = PassTables<Q>::kL1;
	static __host__ __device__ constexpr int offset(int L) {         // float2 offset of the pass of length L
		int o = 0;
		for(int l = kL1; l > L; l /= 8) o += l / 2;                   // (l/8) butterflies x 4 factors
		return o;
	}
};

// Radix-8 pass with the factors W^j..W^4j read from shared memory and W^5j..W^7j formed by one multiplication each.
template <int Q, int L>
__device__ __forceinline__ void pass_radix8_s(float2* __restrict__ T, int t, const float2* __restrict__ tw8) {
	constexpr int s = L / 8, ps = s + s / 8;
	const int blk = t / s, j = t - blk * s;
	const int base = blk * L + j;
	float2* p = T + base + (base >> 3);
	float2 a[8];
#pragma unroll
	for(int m = 0; m < 8; ++m) a[m] = p[m * ps];
	constexpr int kOff = Tw8Tables<Q>::offset(L);
	// table of a pass: [s x (W^j, W^2j)] then [s x (W^3j, W^4j)] (host: make_fft_r8_tables)
	const float4 w12 = *reinterpret_cast<const float4*>(tw8 + kOff + j * 2);
	const float4 w34 = *reinterpret_cast<const float4*>(tw8 + kOff + 2 * s + j * 2);
	dft8(a);
	const float2 w1 = make_float2(w12.x, w12.y), w2 = make_float2(w12.z, w12.w);
	const float2 w3 = make_float2(w34.x, w34.y), w4 = make_float2(w34.z, w34.w);
	a[1] = cmul(a[1], w1);
	a[2] = cmul(a[2], w2);
	a[3] = cmul(a[3], w3);
	a[4] = cmul(a[4], w4);
	a[5] = cmul(a[5], cmul(w4, w1));
	a[6] = cmul(a[6], cmul(w4, w2));
	a[7] = cmul(a[7], cmul(w4, w3));
#pragma unroll
	for(int k = 0; k < 8; ++k) p[k * ps] = a[k];
}

// Frame sample y[m] (m < n = 2M) from the D array of that frame.
__device__ __forceinline__ float frame_from_D(const float* __restrict__ D, int M, int m) {
	if(m < M / 2) return D[m + M / 2];
	if(m < 3 * M / 2) return -D[3 * M / 2 - 1 - m];
	return -D[m - 3 * M / 2];
}

// Window value at index i of a frame of size n whose left/right slopes have lengths `left`/`right`
// (reference: hpp:846-859). slopeL/slopeR = rising slope tables of those lengths.
__device__ __forceinline__ float window_at(int n, int i, int left, int right, const float* __restrict__ slopeL,
                                           const float* __restrict__ slopeR) {
	const int lb = n / 4 - left / 2, rb = n - n / 4 - right / 2;
	if(i < lb) return 0.f;
	if(i < lb + left) return __ldg(&slopeL[i - lb]);
	if(i < rb) return 1.f;
	if(i < rb + right) return __ldg(&slopeR[right - 1 - (i - rb)]);
	return 0.f;
}

}  // namespace pov
