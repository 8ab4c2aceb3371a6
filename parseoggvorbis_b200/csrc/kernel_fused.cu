// Production path: ONE kernel from coded floor posts + residue spectra to PCM.
//
//   floor1 unwrap/render -> nonzero propagate -> inverse coupling -> floor multiply -> inverse MDCT
//   -> window -> overlap-add -> planar/interleaved PCM
// (reference: VorbisStream::parse_audio stages 4.3.2-4.3.7 + VorbisStreamDecodeState, src/ParseOggVorbis.hpp:
//  521-591, 1174-1180, 1213-1268, 1008-1059; the IMDCT contract is src/mdct.h:105.)
//
// Why fused: unfused, every channel-packet of blocksize n moves 2n B of spectrum in, 4n B of frame out, 4n B
// of frame back in and 2n B of PCM out. Here HBM sees the spectrum once (TMA bulk copy into shared memory,
// double buffered) and the PCM once (coalesced 128-bit stores): 4n B per channel-packet = 8 B per PCM sample,
// the algorithmic minimum of SURVEY.md §8(d). Everything in between lives in shared memory and registers.
//
// Work decomposition: one CTA per *run* = up to K consecutive packets of one stream, all channels. The overlap
// of consecutive frames is carried in shared memory (the D array of the previous frame); the first packet of a
// run is a halo that is transformed again instead of exchanged, so runs are independent (no inter-CTA traffic).
// Inside a run the CTA advances in *steps*: one long packet, or a group of consecutive short packets that
// together fill the same buffers (batching by blocksize class).
#include "kernels.h"
#include "fft_core.cuh"
#include "floor_core.cuh"

namespace pov {

// ---- mbarrier / TMA bulk copy (sm_90+ PTX; SASS: UBLKCP + SYNCS) ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra.uni WAIT_DONE;\n\t"
		"bra.uni WAIT_LOOP;\n\t"
		"WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void uncouple_f(float& m, float& a) {   // hpp:1220-1239
	const float mv = m, av = a;
	if(mv > 0.f) {
		if(av > 0.f) a = mv - av;
		else { a = mv; m = mv + av; }
	} else {
		if(av > 0.f) a = mv + av;
		else { a = mv; m = mv - av; }
	}
}

struct StepInfo {
	uint32_t first;      // global index of the step's first packet
	uint32_t count;      // packets in the step
	uint32_t flag;       // blocksize class (0 short / 1 long)
};

struct FusedParams {
	DevBatchView b;
	const DevRun* runs;
	uint32_t floor_cap[2];   // posts capacity (multiple of 4) of a curve block, per blocksize class
	uint32_t scratch_cap;    // posts capacity of the per-warp unwrap scratch (only allocated when > 32)
	uint32_t group_short;    // max short packets per step
	uint32_t slot_floats;    // floats of one spectra buffer = C_max * blocksize1/2
	uint32_t curve_bytes;    // size of the curve-block region
};

// Elementwise stage for C channels (compile time): floor curve evaluation + coupling + floor multiply + DCT-IV
// pre-rotation. Item = (packet g of the step, pair q): complex points j1 = q and j2 = Q-1-q, which together consume
// the four bins 2q, 2q+1, M-2-2q, M-1-2q of every channel (two aligned float2 loads per spectrum).
template <int C>
__device__ __forceinline__ void stage_spectral(const float* __restrict__ raw, unsigned char* __restrict__ curves, uint32_t curve_stride,
                                               uint32_t floor_cap, const float* __restrict__ invdb, float2* __restrict__ T,
                                               int npk, int log2pairs, const float2* __restrict__ rot,
                                               const DevMapping* __restrict__ mp) {
	const int pairs = 1 << log2pairs, Q = 2 * pairs, M = 2 * Q;
	const int tstride = Q + Q / 8;
	const int ncoup = (C > 1) ? (int) mp->n_couplings : 0;
	for(int it = threadIdx.x; it < (npk << log2pairs); it += blockDim.x) {
		const int g = it >> log2pairs, q = it & (pairs - 1);
		const float* R0 = raw + (size_t) (g * C) * M;
		float x0[C], x1[C], x2[C], x3[C];
#pragma unroll
		for(int c = 0; c < C; ++c) {
			const float2 a = *reinterpret_cast<const float2*>(R0 + c * M + 2 * q);
			const float2 d = *reinterpret_cast<const float2*>(R0 + c * M + M - 2 - 2 * q);
			x0[c] = a.x; x1[c] = a.y; x2[c] = d.x; x3[c] = d.y;
		}
		if(C == 2) {
			// two channels can only be coupled with each other: no channel search needed
			for(int k = ncoup - 1; k >= 0; --k) {
				if(mp->coupling_mag[k] == 0) {
					uncouple_f(x0[0], x0[C - 1]); uncouple_f(x1[0], x1[C - 1]); uncouple_f(x2[0], x2[C - 1]); uncouple_f(x3[0], x3[C - 1]);
				} else {
					uncouple_f(x0[C - 1], x0[0]); uncouple_f(x1[C - 1], x1[0]); uncouple_f(x2[C - 1], x2[0]); uncouple_f(x3[C - 1], x3[0]);
				}
			}
		} else if(C > 2) {
			for(int k = ncoup - 1; k >= 0; --k) {
				const int m = mp->coupling_mag[k], a = mp->coupling_ang[k];
				float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
				for(int c = 0; c < C; ++c) {
					if(c == m) { m0 = x0[c]; m1 = x1[c]; m2 = x2[c]; m3 = x3[c]; }
					if(c == a) { a0 = x0[c]; a1 = x1[c]; a2 = x2[c]; a3 = x3[c]; }
				}
				uncouple_f(m0, a0); uncouple_f(m1, a1); uncouple_f(m2, a2); uncouple_f(m3, a3);
#pragma unroll
				for(int c = 0; c < C; ++c) {
					if(c == m) { x0[c] = m0; x1[c] = m1; x2[c] = m2; x3[c] = m3; }
					if(c == a) { x0[c] = a0; x1[c] = a1; x2[c] = a2; x3[c] = a3; }
				}
			}
		}
		const float2 w1 = __ldg(&rot[q]), w2 = __ldg(&rot[Q - 1 - q]);
		float2* T0 = T + (size_t) (g * C) * tstride;
		const int p1 = q + (q >> 3), p2 = (Q - 1 - q) + ((Q - 1 - q) >> 3);
#pragma unroll
		for(int c = 0; c < C; ++c) {
			CurveV3 Cv;
			Cv.bind(curves + (size_t) (g * C + c) * curve_stride, floor_cap);
			float2 fa, fd;
			const uint32_t mode = Cv.hdr[0];
			if(mode == 0) {
				fa = curve_pair(Cv, (uint32_t) (2 * q), invdb);
				fd = curve_pair(Cv, (uint32_t) (M - 2 - 2 * q), invdb);
			} else {
				const float fill = (mode == 1) ? 1.f : 0.f;
				fa = make_float2(fill, fill); fd = fa;
			}
			// hpp:1252 residue *= floor (one rounding each)
			const float y0 = __fmul_rn(x0[c], fa.x), y1 = __fmul_rn(x1[c], fa.y);
			const float y2 = __fmul_rn(x2[c], fd.x), y3 = __fmul_rn(x3[c], fd.y);
			T0[c * tstride + p1] = cmul(make_float2(y0, y3), w1);
			T0[c * tstride + p2] = cmul(make_float2(y2, y1), w2);
		}
	}
}

// Last pass + post-rotation with the D array split in halves: lo = D[0..M/2) (second half of the frame, needed by the
// NEXT packet), hi = D[M/2..M) (first half of the frame, consumed by this packet's overlap-add).
template <int Q>
__device__ __forceinline__ void pass_last_split(const float2* __restrict__ T, int t, const float2* __restrict__ rot,
                                                float* __restrict__ lo, float* __restrict__ hi) {
	constexpr int M = 2 * Q;
	float2 a[8];
	const float2* p = T + 9 * t;
#pragma unroll
	for(int m = 0; m < 8; ++m) a[m] = p[m];
	dft8(a);
	const int k0 = freq_of_pos<Q>(8 * t);           // k0 < Q/8; k = k0 + m*Q/8
	const float2* r = rot + k0;
#pragma unroll
	for(int m = 0; m < 8; ++m) {
		const float2 c = cmul(a[m], __ldg(r + m * (Q / 8)));
		// D[2k] = Re, D[M-1-2k] = -Im;  2k < M/2  <=>  m < 4
		if(m < 4) { lo[2 * k0 + m * (Q / 4)] = c.x; hi[(M / 2 - 1 - 2 * k0) - m * (Q / 4)] = -c.y; }
		else      { hi[2 * k0 + m * (Q / 4) - M / 2] = c.x; lo[(M - 1 - 2 * k0) - m * (Q / 4)] = -c.y; }
	}
}

template <int Q>
__device__ __forceinline__ void stage_fft(float2* T, float* Dlo, float* Dhi, int nf, const float2* rot, const float2* TWP) {
	fft_passes_except_last_p<Q>(T, nf, TWP);
	for(int w = threadIdx.x; w < nf * FftGeom<Q>::kItems; w += blockDim.x) {
		const int f = w / FftGeom<Q>::kItems, t = w - f * FftGeom<Q>::kItems;
		pass_last_split<Q>(T + (size_t) f * FftGeom<Q>::kStride, t, rot, Dlo + (size_t) f * Q, Dhi + (size_t) f * Q);
	}
}

__device__ __forceinline__ float4 rev_neg(float4 v) { return make_float4(-v.w, -v.z, -v.y, -v.x); }
__device__ __forceinline__ float4 rev4(float4 v) { return make_float4(v.w, v.z, v.y, v.x); }
__device__ __forceinline__ float4 neg4(float4 v) { return make_float4(-v.x, -v.y, -v.z, -v.w); }

// Geometry of one emitting packet for the overlap-add stage (all block-uniform).
struct OlaGeom {
	int Hp, H;          // quarter sizes: Hp = n_prev/4, H = n/4 (= length of the lo / hi halves of D)
	int shift;          // index in the current frame of the chunk's first sample: n/4 - n_prev/4
	int lb, lc;         // current frame: left slope begins at lb, has length lc
	int rbp, pr;        // previous frame, relative to its second half: falling slope begins at rbp, has length pr
	const float* slL;   // rising slope table of length lc
	const float* slR;   // rising slope table of length pr (read mirrored)
};

// Four consecutive output samples j..j+3 (j % 4 == 0):
//   out = (0 + prev[n_prev/2 + j] * w_prev) + cur[j + shift] * w_cur        (hpp:1008-1017 in gather form)
// plo = lo half of the previous frame's D, chi = hi half of the current frame's D.
// Every region boundary is a multiple of 16, so the four samples always share one case.
__device__ __forceinline__ float4 ola_quad(const OlaGeom& G, const float* __restrict__ plo, const float* __restrict__ chi, int j) {
	float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
	if(j < G.rbp + G.pr) {                             // previous frame present and its window non-zero (rbp+pr <= 2*Hp)
		// second half of the previous frame: -D[Hp-1-i] for i < Hp, -D[i-Hp] beyond
		const float4 y = (j < G.Hp) ? rev_neg(*reinterpret_cast<const float4*>(plo + G.Hp - 4 - j))
		                            : neg4(*reinterpret_cast<const float4*>(plo + j - G.Hp));
		float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
		if(j >= G.rbp) w = rev4(__ldg(reinterpret_cast<const float4*>(G.slR + G.pr - 4 - (j - G.rbp))));
		acc.x = __fadd_rn(acc.x, __fmul_rn(y.x, w.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(y.y, w.y));
		acc.z = __fadd_rn(acc.z, __fmul_rn(y.z, w.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(y.w, w.w));
	}
	const int ic = j + G.shift;
	if(ic >= G.lb) {                                    // current frame present and its window non-zero (ic < 2H always)
		// first half of the current frame: D[H+i] for i < H, -D[3H-1-i] beyond
		const float4 y = (ic < G.H) ? *reinterpret_cast<const float4*>(chi + ic)
		                            : rev_neg(*reinterpret_cast<const float4*>(chi + 2 * G.H - 4 - ic));
		float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
		if(ic < G.lb + G.lc) w = __ldg(reinterpret_cast<const float4*>(G.slL + ic - G.lb));
		acc.x = __fadd_rn(acc.x, __fmul_rn(y.x, w.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(y.y, w.y));
		acc.z = __fadd_rn(acc.z, __fmul_rn(y.z, w.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(y.w, w.w));
	}
	return acc;
}

// scalar version for ragged tails / unaligned destinations / interleaved output
__device__ __forceinline__ float ola_one(const OlaGeom& G, const float* __restrict__ plo, const float* __restrict__ chi, int j) {
	float acc = 0.f;
	if(j < G.rbp + G.pr) {
		const float y = (j < G.Hp) ? -plo[G.Hp - 1 - j] : -plo[j - G.Hp];
		const float w = (j >= G.rbp) ? __ldg(G.slR + G.pr - 1 - (j - G.rbp)) : 1.f;
		acc = __fadd_rn(acc, __fmul_rn(y, w));
	}
	const int ic = j + G.shift;
	if(ic >= G.lb) {
		const float y = (ic < G.H) ? chi[ic] : -chi[2 * G.H - 1 - ic];
		const float w = (ic < G.lb + G.lc) ? __ldg(G.slL + ic - G.lb) : 1.f;
		acc = __fadd_rn(acc, __fmul_rn(y, w));
	}
	return acc;
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_fused_synth(FusedParams P) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ float s_invdb[256];
	__shared__ uint8_t s_mode_flag[POV_MAX_MODES], s_mode_map[POV_MAX_MODES];

	const DevBatchView& b = P.b;
	const DevRun run = P.runs[blockIdx.x];
	const pov_packet pk0 = b.packets[run.first_packet];
	const pov_stream st = b.streams[pk0.stream];
	const DevSetup* __restrict__ su = &b.setups[st.setup_id];
	const int C = (int) su->channels;
	const int nwarps = kThreads >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t bs0 = su->blocksize[0], bs1 = su->blocksize[1];
	const DevFloor* __restrict__ floors = su->floors;
	const DevMapping* __restrict__ mappings = su->mappings;

	// ---- shared memory carve-up (floats): raw[slot] | T[1.125*slot] | Dlo[2][slot/2] | Dhi[slot/2] | curves | scratch
	const uint32_t slot = P.slot_floats;
	float* raw = reinterpret_cast<float*>(smem);
	float2* T = reinterpret_cast<float2*>(raw + slot);
	float* Dlo = reinterpret_cast<float*>(T) + (size_t) slot + slot / 8;
	float* Dhi = Dlo + slot;
	unsigned char* curves = reinterpret_cast<unsigned char*>(Dhi + slot / 2);
	unsigned char* fscr = curves + P.curve_bytes;    // per-warp unwrap scratch, only when scratch_cap > 32

	for(int i = threadIdx.x; i < 256; i += kThreads) s_invdb[i] = __ldg(&b.inv_db[i]);
	for(int i = threadIdx.x; i < (int) POV_MAX_MODES; i += kThreads) { s_mode_flag[i] = su->mode_blockflag[i]; s_mode_map[i] = su->mode_mapping[i]; }
	if(threadIdx.x == 0) {
		mbar_init(&s_bar, 1);
		mbar_fence_init();
	}
	__syncthreads();

	const uint32_t run_end = run.first_packet + run.n_packets;
	const uint32_t gshort = (bs0 == bs1) ? 1u : P.group_short;

	auto make_step = [&](uint32_t first) {
		StepInfo s;
		s.first = first;
		s.count = 0;
		s.flag = 0;
		if(first >= run_end) return s;
		const uint32_t mode = b.packets[first].mode;
		s.flag = s_mode_flag[mode];
		s.count = 1;
		if(!s.flag && bs0 != bs1)   // group consecutive short packets of the same mode (same mapping, same floors)
			while(s.count < gshort && first + s.count < run_end && b.packets[first + s.count].mode == mode) ++s.count;
		return s;
	};
	auto issue_loads = [&](const StepInfo& s) {     // one elected thread: TMA bulk copies of the step's spectra
		const uint32_t half = (s.flag ? bs1 : bs0) / 2;
		const uint32_t bytes = (uint32_t) C * half * 4u;
		mbar_expect_tx(&s_bar, bytes * s.count);
		for(uint32_t g = 0; g < s.count; ++g)
			tma_bulk_g2s(raw + (size_t) g * C * half, b.spectra + b.spec_off[s.first + g], bytes, &s_bar);
	};

	StepInfo cur = make_step(run.first_packet);
	if(threadIdx.x == 0 && cur.count) issue_loads(cur);
	uint32_t phase = 0;
	// overlap carried from the previous packet: the lo half of its D array and its geometry
	int prev_valid = 0, prev_n = 0, prev_right = 0;
	const float* prev_lo = nullptr;    // channel 0; channels are prev_n/4 apart
	int step_idx = 0;

	while(cur.count) {
		const int buf = step_idx & 1;
		const StepInfo nxt = make_step(cur.first + cur.count);
		const uint32_t flag = cur.flag;
		const int n = (int) (flag ? bs1 : bs0), M = n / 2, Q = n / 4;
		const int log2Q = 31 - __clz(Q);
		const int npk = (int) cur.count, nf = npk * C;
		const uint32_t mode0 = b.packets[cur.first].mode;
		const DevMapping* mp = &mappings[s_mode_map[mode0]];
		float* Dlo_cur = Dlo + (size_t) buf * (slot / 2);
		const uint32_t cells = (uint32_t) M / 4;
		const uint32_t fcap = P.floor_cap[flag];
		const uint32_t curve_stride = CurveV3::bytes(fcap, cells);

		// ---- stage 1: floor unwrap + curve records, one warp per (packet, channel) curve ----
		for(int f = warp; f < nf; f += nwarps) {
			const int g = f / C, c = f - g * C;
			const uint32_t p = cur.first + g;
			const pov_packet pk = b.packets[p];
			CurveV3 Cv;
			Cv.bind(curves + (size_t) f * curve_stride, fcap);
			const uint32_t used = pk.floor_used;
			if(!((used >> c) & 1)) {
				// no curve decoded: the reference multiplies by its zero-initialised floor buffer if the channel became
				// "used" through coupling (hpp:1159,1247) and leaves the residue untouched otherwise
				uint32_t prop = used;
				for(uint32_t k = 0; k < mp->n_couplings; ++k) {   // hpp:1174-1180
					const uint32_t m = mp->coupling_mag[k], a = mp->coupling_ang[k];
					if(((prop >> m) | (prop >> a)) & 1) prop |= (1u << m) | (1u << a);
				}
				if(lane == 0) Cv.hdr[0] = ((prop >> c) & 1) ? 2u : 1u;
				continue;
			}
			const DevFloor* F = &floors[mp->floor_of_ch[c]];
			uint64_t yo = pk.ys_off;
			for(int cc = 0; cc < c; ++cc)
				if((used >> cc) & 1) yo += floors[mp->floor_of_ch[cc]].n_posts;
			uint32_t stt;
			if(F->n_posts <= 32) {
				stt = floor1_curve_warp32(F, b.ys + yo, Cv, cells, (uint32_t) n, lane);
			} else {
				FloorScratch W;
				W.bind(fscr + (size_t) warp * floor_scratch_stride(P.scratch_cap), P.scratch_cap);
				stt = floor1_unwrap_warp(F, b.ys + yo, W, lane);
				stt |= floor1_range_check_warp(W, (uint32_t) n, lane);
				const uint32_t ns = *W.nseg;
				if(lane == 0) Cv.hdr[0] = 0;
				for(uint32_t base = 0; base < ns; base += 32) {
					const uint32_t s = base + lane;
					const bool have = s < ns;
					const uint32_t x0 = have ? W.segx[s] : 0u, y0 = have ? W.segy[s] : 0u;
					const uint32_t x1 = (s + 1 < ns) ? W.segx[s + 1] : 0u, y1 = (s + 1 < ns) ? W.segy[s + 1] : 0u;
					curve_build_warp(Cv, ns, x0, y0, x1, y1, have, base, cells, lane, base == 0);
				}
				curve_scan_cells_warp(Cv, cells, lane);
			}
			if(stt && lane == 0) atomicOr(&b.status[p], stt);
		}
		// ---- wait for this step's spectra (TMA), then everyone sees curves + raw ----
		mbar_wait(&s_bar, phase);
		phase ^= 1;
		__syncthreads();

		// ---- stage 2: floor evaluation + coupling + floor multiply + pre-rotation -> T ----
		{
			const float2* rot = su->rot[flag];
			const int lp = log2Q - 1;
			switch(C) {
				case 1: stage_spectral<1>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				case 2: stage_spectral<2>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				case 3: stage_spectral<3>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				case 4: stage_spectral<4>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				case 5: stage_spectral<5>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				case 6: stage_spectral<6>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				case 7: stage_spectral<7>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
				default: stage_spectral<8>(raw, curves, curve_stride, fcap, s_invdb, T, npk, lp, rot, mp); break;
			}
		}
		__syncthreads();
		// the spectra buffer and the curve blocks are free again: prefetch the next step's spectra behind stages 3-4
		if(threadIdx.x == 0 && nxt.count) issue_loads(nxt);

		// ---- stage 3: FFT passes + post-rotation -> D (lo / hi halves) ----
		{
			const float2* rot = su->rot[flag];
			const float2* TWP = su->fftp[flag];
			switch(Q) {
				case 16:   stage_fft<16>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				case 32:   stage_fft<32>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				case 64:   stage_fft<64>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				case 128:  stage_fft<128>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				case 256:  stage_fft<256>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				case 512:  stage_fft<512>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				case 1024: stage_fft<1024>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
				default:   stage_fft<2048>(T, Dlo_cur, Dhi, nf, rot, TWP); break;
			}
		}
		__syncthreads();

		// ---- stage 4: window + overlap-add + emit (hpp:1008-1059 in gather form) ----
		for(int g = 0; g < npk; ++g) {
			const uint32_t p = cur.first + g;
			const pov_packet pk = b.packets[p];
			// hpp:844-847: short blocks always use blocksize0 slopes; long blocks follow their own prev/next flags
			const int lc = (int) ((flag && (pk.window_flags & 1)) ? bs1 : bs0) / 2;
			const int rc = (int) ((flag && (pk.window_flags & 2)) ? bs1 : bs0) / 2;
			const float* cur_lo = Dlo_cur + (size_t) g * C * Q;
			const float* cur_hi = Dhi + (size_t) g * C * Q;
			const bool emits = prev_valid && pk.emit_frames > 0 && p != st.first_packet;
			if(emits) {
				OlaGeom G;
				G.Hp = prev_n / 4; G.H = Q;
				G.shift = Q - prev_n / 4;
				G.lc = lc; G.lb = Q - lc / 2;
				G.pr = prev_right; G.rbp = prev_n / 4 - prev_right / 2;
				G.slL = su->slope[lc == (int) bs1 / 2 ? 1 : 0];
				G.slR = su->slope[prev_right == (int) bs1 / 2 ? 1 : 0];
				const uint32_t emit = pk.emit_frames;
				const bool planar = (b.pcm_layout == POV_PCM_PLANAR);
				const uint64_t chan_base = st.pcm_base + pk.pcm_off;
				if(planar && (emit & 3u) == 0 && ((chan_base | st.pcm_frames) & 3ull) == 0) {
					// 128-bit path: one thread per 4 consecutive frames of one channel
					const uint32_t quads = emit >> 2;
					for(int c = 0; c < C; ++c) {
						const float* plo = prev_lo + (size_t) c * G.Hp;
						const float* chi = cur_hi + (size_t) c * Q;
						float* dst = b.pcm + chan_base + (uint64_t) c * st.pcm_frames;
						for(uint32_t qd = threadIdx.x; qd < quads; qd += kThreads)
							*reinterpret_cast<float4*>(dst + 4 * qd) = ola_quad(G, plo, chi, (int) (4 * qd));
					}
				} else {
					const uint32_t total = emit * (uint32_t) C;
					for(uint32_t e = threadIdx.x; e < total; e += kThreads) {
						uint32_t c, j;
						if(planar) { c = e / emit; j = e - c * emit; } else { j = e / (uint32_t) C; c = e - j * (uint32_t) C; }
						const float v = ola_one(G, prev_lo + (size_t) c * G.Hp, cur_hi + (size_t) c * Q, (int) j);
						const uint64_t fidx = pk.pcm_off + j;
						const uint64_t o = planar ? st.pcm_base + (uint64_t) c * st.pcm_frames + fidx : st.pcm_base + fidx * (uint64_t) C + c;
						b.pcm[o] = v;
					}
				}
			}
			prev_valid = 1; prev_n = n; prev_right = rc; prev_lo = cur_lo;
		}
		// No barrier needed here: the next step writes curves/T before its own barriers; Dhi and Dlo[buf^1] are not
		// written again before two more barriers.
		cur = nxt;
		++step_idx;
	}
}

static uint32_t fused_threads(uint32_t max_channels, uint32_t max_blocksize) {
	// one 8-point work item per thread per pass for the long block: C * n/32 threads -> 128, 256 or 512
	const uint32_t want = max_channels * (max_blocksize / 32);
	if(want <= 128) return 128;
	if(want <= 256) return 256;
	return 512;
}

static void fused_layout(uint32_t max_channels, uint32_t max_blocksize, uint32_t min_blocksize, const uint32_t floor_cap[2],
                         FusedParams& P, uint32_t& threads, size_t& smem) {
	threads = fused_threads(max_channels, max_blocksize);
	uint32_t group = max_blocksize / min_blocksize;
	if(group > 8) group = 8;
	if(group < 1) group = 1;
	P.group_short = group;
	P.floor_cap[0] = floor_cap[0]; P.floor_cap[1] = floor_cap[1];
	P.scratch_cap = floor_cap[0] > floor_cap[1] ? floor_cap[0] : floor_cap[1];
	P.slot_floats = max_channels * (max_blocksize / 2);
	const size_t long_bytes = (size_t) max_channels * CurveV3::bytes(floor_cap[1], max_blocksize / 8);
	const size_t short_bytes = (size_t) group * max_channels * CurveV3::bytes(floor_cap[0], min_blocksize / 8);
	// a setup with blocksize0 == blocksize1 runs every packet as a "short" step of one packet with the long geometry
	const size_t same_bytes = (size_t) max_channels * CurveV3::bytes(P.scratch_cap, max_blocksize / 8);
	size_t cb = long_bytes > short_bytes ? long_bytes : short_bytes;
	if(same_bytes > cb) cb = same_bytes;
	P.curve_bytes = (uint32_t) ((cb + 15) & ~(size_t) 15);
	const size_t slot = P.slot_floats;
	const size_t floats = slot + (slot + slot / 8) + slot + slot / 2;       // raw | T | Dlo[2] | Dhi
	smem = floats * sizeof(float) + P.curve_bytes + 128;
	if(P.scratch_cap > 32) smem += (size_t) (threads / 32) * floor_scratch_stride(P.scratch_cap);
}

size_t fused_smem_bytes(uint32_t max_channels, uint32_t max_blocksize, uint32_t min_blocksize, const uint32_t floor_cap[2]) {
	FusedParams P;
	uint32_t threads;
	size_t smem;
	fused_layout(max_channels, max_blocksize, min_blocksize, floor_cap, P, threads, smem);
	return smem;
}

cudaError_t launch_fused(const DevBatchView& b, const DevRun* runs, uint32_t n_runs, uint32_t max_channels,
                         uint32_t max_blocksize, uint32_t min_blocksize, const uint32_t floor_cap[2], cudaStream_t st, uint64_t* launches) {
	if(n_runs == 0) return cudaSuccess;
	FusedParams P;
	uint32_t threads;
	size_t smem;
	fused_layout(max_channels, max_blocksize, min_blocksize, floor_cap, P, threads, smem);
	P.b = b;
	P.runs = runs;
	if(smem > 227 * 1024) return cudaErrorInvalidConfiguration;
	cudaError_t e;
	if(threads == 128) {
		e = cudaFuncSetAttribute(k_fused_synth<128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if(e != cudaSuccess) return e;
		k_fused_synth<128, 6><<<n_runs, 128, smem, st>>>(P);
	} else if(threads == 256) {
		e = cudaFuncSetAttribute(k_fused_synth<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if(e != cudaSuccess) return e;
		k_fused_synth<256, 3><<<n_runs, 256, smem, st>>>(P);
	} else {
		e = cudaFuncSetAttribute(k_fused_synth<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
		if(e != cudaSuccess) return e;
		k_fused_synth<512, 1><<<n_runs, 512, smem, st>>>(P);
	}
	if(launches) ++*launches;
	return cudaGetLastError();
}

}  // namespace pov
