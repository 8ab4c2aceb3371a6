// Production path: ONE kernel from coded floor posts + residue spectra to PCM.
//
//   floor1 unwrap/render -> nonzero propagate -> inverse coupling -> floor multiply -> inverse MDCT
//   -> window -> overlap-add -> planar/interleaved PCM
// (reference: VorbisStream::parse_audio stages 4.3.2-4.3.7 + VorbisStreamDecodeState, src/ParseOggVorbis.hpp:
//  521-591, 1174-1180, 1213-1268, 1008-1059; the IMDCT contract is src/mdct.h:105.)
//
// Why fused: unfused, every channel-packet of blocksize n moves 2n B of spectrum in, 4n B of frame out, 4n B
// of frame back in and 2n B of PCM out. Here HBM sees the spectrum once (TMA bulk copy into shared memory,
// double buffered) and the PCM once (coalesced stores): 4n B per channel-packet = 8 B per PCM sample, the
// algorithmic minimum of SURVEY.md §8(d). Everything in between lives in shared memory and registers.
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
	uint32_t floor_cap;      // posts capacity of the per-warp floor scratch (multiple of 4)
	uint32_t group_short;    // max short packets per step
	uint32_t slot_floats;    // floats of one raw/floor buffer set = C_max * blocksize1/2
};

// Elementwise stage for C channels (compile time): coupling + floor multiply + DCT-IV pre-rotation.
// Item = (packet g of the step, pair q): complex points j1 = q and j2 = Q-1-q, which together consume the four
// bins 2q, 2q+1, M-2-2q, M-1-2q of every channel (two aligned float2 loads per array).
template <int C>
__device__ __forceinline__ void stage_spectral(const float* __restrict__ raw, const float* __restrict__ flo, float2* __restrict__ T,
                                               int npk, int Q, int tstride, const float2* __restrict__ rot,
                                               const DevMapping* __restrict__ mp) {
	const int M = 2 * Q, pairs = Q / 2;
	const int ncoup = (int) mp->n_couplings;
	for(int it = threadIdx.x; it < npk * pairs; it += blockDim.x) {
		const int g = it / pairs, q = it - g * pairs;
		float x0[C], x1[C], x2[C], x3[C];
#pragma unroll
		for(int c = 0; c < C; ++c) {
			const float* R = raw + (size_t) (g * C + c) * M;
			const float2 a = *reinterpret_cast<const float2*>(R + 2 * q);
			const float2 d = *reinterpret_cast<const float2*>(R + M - 2 - 2 * q);
			x0[c] = a.x; x1[c] = a.y; x2[c] = d.x; x3[c] = d.y;
		}
		if(C > 1) {
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
#pragma unroll
		for(int c = 0; c < C; ++c) {
			const float* F = flo + (size_t) (g * C + c) * M;
			const float2 fa = *reinterpret_cast<const float2*>(F + 2 * q);
			const float2 fd = *reinterpret_cast<const float2*>(F + M - 2 - 2 * q);
			// hpp:1252 residue *= floor (one rounding each; F holds 1.0 / 0.0 for channels without a curve)
			const float y0 = __fmul_rn(x0[c], fa.x), y1 = __fmul_rn(x1[c], fa.y);
			const float y2 = __fmul_rn(x2[c], fd.x), y3 = __fmul_rn(x3[c], fd.y);
			float2* Tf = T + (size_t) (g * C + c) * tstride;
			Tf[tpad(q)] = cmul(make_float2(y0, y3), w1);
			Tf[tpad(Q - 1 - q)] = cmul(make_float2(y2, y1), w2);
		}
	}
}

template <int Q>
__device__ __forceinline__ void stage_fft(float2* T, float* D, int nf, const float2* rot, const float2* W) {
	fft_passes_except_last<Q>(T, nf, W);
	for(int w = threadIdx.x; w < nf * FftGeom<Q>::kItems; w += blockDim.x) {
		const int f = w / FftGeom<Q>::kItems, t = w - f * FftGeom<Q>::kItems;
		pass_last_to_D<Q>(T + (size_t) f * FftGeom<Q>::kStride, t, rot, D + (size_t) f * 2 * Q);
	}
}

__device__ __forceinline__ void slope_lengths_f(const DevSetup& su, uint32_t flag, uint32_t wflags, int& left, int& right) {
	left = (int) ((flag && (wflags & 1)) ? su.blocksize[1] : su.blocksize[0]) / 2;    // hpp:844-847
	right = (int) ((flag && (wflags & 2)) ? su.blocksize[1] : su.blocksize[0]) / 2;
}

__global__ void __launch_bounds__(512) k_fused_synth(FusedParams P) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t s_bar[2];
	__shared__ float s_invdb[256];

	const DevBatchView& b = P.b;
	const DevRun run = P.runs[blockIdx.x];
	const pov_packet pk0 = b.packets[run.first_packet];
	const pov_stream st = b.streams[pk0.stream];
	const DevSetup& su = b.setups[st.setup_id];
	const int C = (int) su.channels;
	const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	// ---- shared memory carve-up (floats): raw[2][slot] | floor[slot] | T[1.125*slot] | D[2][slot] | floor scratch
	const uint32_t slot = P.slot_floats;
	float* raw = reinterpret_cast<float*>(smem);
	float* flo = raw + 2 * (size_t) slot;
	float2* T = reinterpret_cast<float2*>(flo + slot);
	float* D = reinterpret_cast<float*>(T) + (size_t) slot + slot / 8;
	unsigned char* fscr = reinterpret_cast<unsigned char*>(D + 2 * (size_t) slot);
	FloorScratch S;
	S.bind(fscr + (size_t) warp * floor_scratch_stride(P.floor_cap), P.floor_cap);

	for(int i = threadIdx.x; i < 256; i += blockDim.x) s_invdb[i] = __ldg(&b.inv_db[i]);
	if(threadIdx.x == 0) {
		mbar_init(&s_bar[0], 1);
		mbar_init(&s_bar[1], 1);
		mbar_fence_init();
	}
	__syncthreads();

	const uint32_t run_end = run.first_packet + run.n_packets;
	const uint32_t bs0 = su.blocksize[0], bs1 = su.blocksize[1];
	const uint32_t gshort = (bs0 == bs1) ? 1u : P.group_short;

	auto make_step = [&](uint32_t first) {
		StepInfo s;
		s.first = first;
		s.count = 0;
		s.flag = 0;
		if(first >= run_end) return s;
		const uint32_t mode = b.packets[first].mode;
		s.flag = su.mode_blockflag[mode];
		s.count = 1;
		if(!s.flag && bs0 != bs1)   // group consecutive short packets of the same mode (same mapping, same floors)
			while(s.count < gshort && first + s.count < run_end && b.packets[first + s.count].mode == mode) ++s.count;
		return s;
	};
	auto issue_loads = [&](const StepInfo& s, int buf) {     // one elected thread: TMA bulk copies of the step's spectra
		const uint32_t half = su.blocksize[s.flag] / 2;
		const uint32_t bytes = (uint32_t) C * half * 4u;
		mbar_expect_tx(&s_bar[buf], bytes * s.count);
		for(uint32_t g = 0; g < s.count; ++g)
			tma_bulk_g2s(raw + (size_t) buf * slot + (size_t) g * C * half, b.spectra + b.spec_off[s.first + g], bytes, &s_bar[buf]);
	};

	StepInfo cur = make_step(run.first_packet);
	if(threadIdx.x == 0 && cur.count) issue_loads(cur, 0);
	uint32_t phase[2] = {0, 0};
	// overlap carried from the previous step: last packet's D array and geometry
	int prev_valid = 0, prev_n = 0, prev_right = 0;
	const float* prevD = nullptr;    // D of the previous packet, channel 0 (channels are prev_n/2 apart)
	int step_idx = 0;

	while(cur.count) {
		const int buf = step_idx & 1;
		const StepInfo nxt = make_step(cur.first + cur.count);
		if(threadIdx.x == 0 && nxt.count) issue_loads(nxt, buf ^ 1);   // prefetch: raw[buf^1] was consumed a step ago

		const uint32_t flag = cur.flag;
		const int n = (int) su.blocksize[flag], M = n / 2, Q = n / 4;
		const int npk = (int) cur.count, nf = npk * C;
		const DevMapping* mp = &su.mappings[su.mode_mapping[b.packets[cur.first].mode]];
		float* rawb = raw + (size_t) buf * slot;
		float* Dcur = D + (size_t) buf * slot;

		// ---- stage 1: floor curves, one warp per (packet, channel) [x bin slices when warps outnumber curves] ----
		{
			const int nsl = (nf < nwarps) ? nwarps / nf : 1;
			for(int w = warp; w < nf * nsl; w += nwarps) {
				const int f = w / nsl, sl = w - f * nsl;
				const int g = f / C, c = f - g * C;
				const uint32_t p = cur.first + g;
				const pov_packet pk = b.packets[p];
				const DevMapping* mpp = &su.mappings[su.mode_mapping[pk.mode]];
				uint32_t used = pk.floor_used, prop = used;
				for(uint32_t k = 0; k < mpp->n_couplings; ++k) {   // hpp:1174-1180
					const uint32_t m = mpp->coupling_mag[k], a = mpp->coupling_ang[k];
					if(((prop >> m) | (prop >> a)) & 1) prop |= (1u << m) | (1u << a);
				}
				float* Fo = flo + (size_t) f * M;
				const int b0 = (M * sl) / nsl, b1 = (M * (sl + 1)) / nsl;
				if(!((used >> c) & 1)) {
					// no curve decoded: the reference multiplies by its zero-initialised floor buffer if the channel
					// became "used" through coupling (hpp:1159,1247), and leaves the residue untouched otherwise
					const float fill = ((prop >> c) & 1) ? 0.f : 1.f;
					for(int x = b0 + lane; x < b1; x += 32) Fo[x] = fill;
					continue;
				}
				const DevFloor* F = &su.floors[mpp->floor_of_ch[c]];
				uint64_t yo = pk.ys_off;
				for(int cc = 0; cc < c; ++cc)
					if((used >> cc) & 1) yo += su.floors[mpp->floor_of_ch[cc]].n_posts;
				uint32_t stt = floor1_unwrap_warp(F, b.ys + yo, S, lane);
				if(sl == 0) {
					stt |= floor1_range_check_warp(S, (uint32_t) n, lane);
					if(stt && lane == 0) atomicOr(&b.status[p], stt);
				}
				floor1_render_warp(S, (uint32_t) b0, (uint32_t) b1, lane, [&](uint32_t x, uint32_t y) { Fo[x] = s_invdb[y & 255]; });
			}
		}
		// ---- wait for this step's spectra (TMA), then everyone sees floor + raw ----
		mbar_wait(&s_bar[buf], phase[buf]);
		phase[buf] ^= 1;
		__syncthreads();

		// ---- stage 2: coupling + floor multiply + pre-rotation -> T ----
		{
			const int tstride = Q + Q / 8;
			const float2* rot = su.rot[flag];
			switch(C) {
				case 1: stage_spectral<1>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				case 2: stage_spectral<2>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				case 3: stage_spectral<3>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				case 4: stage_spectral<4>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				case 5: stage_spectral<5>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				case 6: stage_spectral<6>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				case 7: stage_spectral<7>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
				default: stage_spectral<8>(rawb, flo, T, npk, Q, tstride, rot, mp); break;
			}
		}
		__syncthreads();

		// ---- stage 3: FFT passes + post-rotation -> D ----
		{
			const float2* rot = su.rot[flag];
			const float2* W = su.fft[flag];
			switch(Q) {
				case 16:   stage_fft<16>(T, Dcur, nf, rot, W); break;
				case 32:   stage_fft<32>(T, Dcur, nf, rot, W); break;
				case 64:   stage_fft<64>(T, Dcur, nf, rot, W); break;
				case 128:  stage_fft<128>(T, Dcur, nf, rot, W); break;
				case 256:  stage_fft<256>(T, Dcur, nf, rot, W); break;
				case 512:  stage_fft<512>(T, Dcur, nf, rot, W); break;
				case 1024: stage_fft<1024>(T, Dcur, nf, rot, W); break;
				default:   stage_fft<2048>(T, Dcur, nf, rot, W); break;
			}
		}
		__syncthreads();

		// ---- stage 4: window + overlap-add + emit (hpp:1008-1059 in gather form) ----
		for(int g = 0; g < npk; ++g) {
			const uint32_t p = cur.first + g;
			const pov_packet pk = b.packets[p];
			int lc, rc;
			slope_lengths_f(su, flag, pk.window_flags, lc, rc);
			const float* Dc = Dcur + (size_t) g * C * M;
			const bool emits = prev_valid && pk.emit_frames > 0 && p != st.first_packet && !(run.halo && p == run.first_packet);
			if(emits) {
				const int np = prev_n, Mp = np / 2;
				const float* slL = su.slope[lc == (int) bs1 / 2 ? 1 : 0];
				const float* slR = su.slope[prev_right == (int) bs1 / 2 ? 1 : 0];
				const int shift = n / 4 - np / 4;
				const int lb = n / 4 - lc / 2;
				const int rb = np - np / 4 - prev_right / 2;
				const uint32_t emit = pk.emit_frames;
				const uint32_t total = emit * (uint32_t) C;
				const bool planar = (b.pcm_layout == POV_PCM_PLANAR);
				for(uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
					uint32_t c, j;
					if(planar) { c = e / emit; j = e - c * emit; } else { j = e / (uint32_t) C; c = e - j * (uint32_t) C; }
					float acc = 0.f;
					const int ip = Mp + (int) j;
					if(ip < np) {
						float wv;
						if(ip < rb) wv = 1.f;
						else if(ip < rb + prev_right) wv = __ldg(&slR[prev_right - 1 - (ip - rb)]);
						else wv = 0.f;
						acc = __fadd_rn(acc, __fmul_rn(frame_from_D(prevD + (size_t) c * Mp, Mp, ip), wv));
					}
					const int ic = (int) j + shift;
					if(ic >= 0 && ic < n) {
						float wv;
						if(ic < lb) wv = 0.f;
						else if(ic < lb + lc) wv = __ldg(&slL[ic - lb]);
						else wv = 1.f;      // ic < M always: the right slope of the current frame is never reached here
						acc = __fadd_rn(acc, __fmul_rn(frame_from_D(Dc + (size_t) c * M, M, ic), wv));
					}
					const uint64_t fidx = pk.pcm_off + j;
					const uint64_t o = planar ? st.pcm_base + (uint64_t) c * st.pcm_frames + fidx : st.pcm_base + fidx * (uint64_t) C + c;
					b.pcm[o] = acc;
				}
			}
			prev_valid = 1; prev_n = n; prev_right = rc; prevD = Dc;
		}
		// No barrier needed here: the next step only writes flo/T/raw before its own barriers, D[buf^1] not before
		// three barriers from now.
		cur = nxt;
		++step_idx;
	}
}

cudaError_t launch_fused(const DevBatchView& b, const DevRun* runs, uint32_t n_runs, uint32_t max_channels,
                         uint32_t max_blocksize, uint32_t min_blocksize, uint32_t floor_cap, cudaStream_t st, uint64_t* launches) {
	if(n_runs == 0) return cudaSuccess;
	FusedParams P;
	P.b = b;
	P.runs = runs;
	P.floor_cap = floor_cap;
	P.group_short = max_blocksize / min_blocksize;
	if(P.group_short > 8) P.group_short = 8;
	if(P.group_short < 1) P.group_short = 1;
	P.slot_floats = max_channels * (max_blocksize / 2);
	// one 8-point work item per thread per pass for the long block: C * n/32 threads, within [128, 512]
	uint32_t threads = max_channels * (max_blocksize / 32);
	threads = (threads + 31u) & ~31u;
	if(threads < 128) threads = 128;
	if(threads > 512) threads = 512;
	const size_t floats = (size_t) P.slot_floats * 2 + P.slot_floats + (P.slot_floats + P.slot_floats / 8) + (size_t) P.slot_floats * 2;
	const size_t smem = floats * sizeof(float) + (size_t) (threads / 32) * floor_scratch_stride(floor_cap) + 128;
	if(smem > 227 * 1024) return cudaErrorInvalidConfiguration;
	cudaError_t e = cudaFuncSetAttribute(k_fused_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if(e != cudaSuccess) return e;
	k_fused_synth<<<n_runs, threads, smem, st>>>(P);
	if(launches) ++*launches;
	return cudaGetLastError();
}

size_t fused_smem_bytes(uint32_t max_channels, uint32_t max_blocksize, uint32_t floor_cap) {
	uint32_t threads = max_channels * (max_blocksize / 32);
	threads = (threads + 31u) & ~31u;
	if(threads < 128) threads = 128;
	if(threads > 512) threads = 512;
	const size_t slot = (size_t) max_channels * (max_blocksize / 2);
	const size_t floats = slot * 2 + slot + (slot + slot / 8) + slot * 2;
	return floats * sizeof(float) + (size_t) (threads / 32) * floor_scratch_stride(floor_cap) + 128;
}

}  // namespace pov
