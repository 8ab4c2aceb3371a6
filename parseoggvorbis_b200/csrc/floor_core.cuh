// floor1 curve synthesis for one channel-packet by one warp (sm_100a).
//
// Behaviour follows the reference (paths relative to its root):
//   step 1, amplitude unwrap      src/ParseOggVorbis.hpp:521-559, render_point src/Utils.hpp:122-137
//   step 2, line rasterisation    src/ParseOggVorbis.hpp:563-585, render_line  src/Utils.hpp:143-183
//   dB lookup                     src/ParseOggVorbis.hpp:586-589, src/inverse_db_table.h:13-78
// but not its structure: the neighbour search is precomputed per setup (it depends on the X list only), posts are
// processed level by level of the neighbour DAG with one lane per post, flagged posts are compacted with
// ballot/popc prefix sums, and every bin is evaluated in closed form  y = y0 +/- floor((x-x0)*|dy| / dx)
// (identical to the reference's error-accumulating loop) so that lanes are independent.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pov_internal.h"

namespace pov {

// Per-warp scratch in shared memory, sized for `cap` posts (cap = setup's largest post count, rounded up).
struct FloorScratch {
	uint32_t* fy;      // [cap]   coded value on entry, final Y after unwrap
	uint16_t* segx;    // [cap]   flagged posts in ascending x
	uint16_t* segy;    // [cap]   final Y * multiplier (clamped to 0xFFFF)
	uint8_t*  flag;    // [cap]   step2_flag
	uint32_t* nseg;    // [1]
	static __host__ __device__ constexpr uint32_t bytes(uint32_t cap) { return cap * 4 + cap * 2 + cap * 2 + cap + 4 + 12; }
	__device__ __forceinline__ void bind(unsigned char* base, uint32_t cap) {   // base 16-byte aligned, cap % 4 == 0
		fy = reinterpret_cast<uint32_t*>(base);
		segx = reinterpret_cast<uint16_t*>(base + cap * 4);
		segy = reinterpret_cast<uint16_t*>(base + cap * 6);
		flag = base + cap * 8;
		nseg = reinterpret_cast<uint32_t*>(base + cap * 9);
	}
};
__host__ __device__ constexpr uint32_t floor_scratch_stride(uint32_t cap) { return (FloorScratch::bytes(cap) + 15u) & ~15u; }

// Unwrap + compaction. All 32 lanes of the warp must call. Returns POV_PKT_* status bits (warp-uniform).
__device__ __forceinline__ uint32_t floor1_unwrap_warp(const DevFloor* __restrict__ F, const uint16_t* __restrict__ ys,
                                                       FloorScratch& S, int lane) {
	const int posts = F->n_posts;
	const uint32_t range = F->range;
	uint32_t bad = 0;
	__syncwarp();
	for(int i = lane; i < posts; i += 32) {
		S.fy[i] = ys[i];
		S.flag[i] = (i < 2) ? 1 : 0;
	}
	__syncwarp();
	const int levels = F->n_levels;
	for(int lv = 1; lv < levels; ++lv) {
		for(int i = lane; i < posts; i += 32) {
			if(F->level[i] != lv) continue;
			const int l = F->lo[i], h = F->hi[i];
			const uint32_t x0 = F->xs[l], x1 = F->xs[h], X = F->xs[i];
			const uint32_t y0 = S.fy[l], y1 = S.fy[h];
			const uint32_t adx = x1 - x0;
			const bool up = y1 >= y0;
			const uint32_t ady = up ? (y1 - y0) : (y0 - y1);
			const uint32_t off = (ady * (X - x0)) / adx;
			const uint32_t predicted = up ? y0 + off : y0 - off;
			const uint32_t val = S.fy[i];
			if(predicted > range) bad |= POV_PKT_FLOOR_PREDICTED;
			const uint32_t high_room = range - predicted, low_room = predicted;
			const uint32_t room = min(high_room, low_room) * 2;
			uint32_t fin = predicted;
			if(val != 0) {
				S.flag[l] = 1; S.flag[h] = 1; S.flag[i] = 1;
				if(val >= room) fin = (high_room > low_room) ? val - low_room + predicted : predicted - val + high_room - 1;
				else fin = (val & 1) ? predicted - (val + 1) / 2 : predicted + val / 2;
			}
			S.fy[i] = fin;
		}
		__syncwarp();
	}
	// flagged posts in ascending-x order
	const uint32_t mult = F->multiplier;
	uint32_t cnt = 0;
	for(int base = 0; base < posts; base += 32) {
		const int s = base + lane;
		const bool valid = s < posts;
		const int i = valid ? F->sorted_idx[s] : 0;
		const bool f = valid && S.flag[i];
		const uint32_t mask = __ballot_sync(0xffffffffu, f);
		if(f) {
			const uint32_t pos = cnt + __popc(mask & ((1u << lane) - 1u));
			S.segx[pos] = F->xs[i];
			S.segy[pos] = (uint16_t) min(S.fy[i] * mult, 0xFFFFu);
		}
		cnt += __popc(mask);
	}
	if(lane == 0) *S.nseg = cnt;
	__syncwarp();
	return __reduce_or_sync(0xffffffffu, bad);
}

// Value of the rendered curve at bin x inside segment [x0,x1) from y0 to y1 (closed form of Utils.hpp:143-183).
__device__ __forceinline__ uint32_t floor1_line_at(uint32_t x0, uint32_t y0, uint32_t adx, uint32_t ady, bool up,
                                                   float rinv, uint32_t x) {
	const uint32_t e = (x - x0) * ady;
	int q = (int) ((float) e * rinv);
	const int r = (int) e - q * (int) adx;
	if(r < 0) --q; else if(r >= (int) adx) ++q;
	return up ? y0 + (uint32_t) q : y0 - (uint32_t) q;
}

// hpp:587 CHECK(floor[i] < 256) over the n bins the reference renders, evaluated from the segment end points
// (each segment is monotone, so its maximum over [0,n) is at a rendered end point). Warp-uniform result.
__device__ __forceinline__ uint32_t floor1_range_check_warp(const FloorScratch& S, uint32_t n, int lane) {
	const int nseg = *S.nseg;
	bool bad = false;
	for(int s = lane; s < nseg; s += 32) {
		const uint32_t x0 = S.segx[s], y0 = S.segy[s];
		if(x0 < n && y0 >= 256) bad = true;
		if(s + 1 < nseg) {
			const uint32_t x1 = S.segx[s + 1], y1 = S.segy[s + 1];
			if(x0 < n && x1 > n - 1 && x1 > x0) {   // the segment is clipped at bin n-1
				const uint32_t adx = x1 - x0;
				const bool up = y1 >= y0;
				const uint32_t ady = up ? y1 - y0 : y0 - y1;
				if(floor1_line_at(x0, y0, adx, ady, up, 1.0f / (float) adx, n - 1) >= 256) bad = true;
			}
		}
	}
	return __any_sync(0xffffffffu, bad) ? (uint32_t) POV_PKT_FLOOR_RANGE : 0u;
}

// Render bins [b0,b1) of the curve. Sink is called as sink(x, y) for every bin exactly once.
template <class Sink>
__device__ __forceinline__ void floor1_render_warp(const FloorScratch& S, uint32_t b0, uint32_t b1, int lane, Sink sink) {
	const int nseg = *S.nseg;
	for(int s = 0; s + 1 < nseg; ++s) {
		const uint32_t x0 = S.segx[s], x1 = S.segx[s + 1];
		if(x1 <= b0) continue;
		if(x0 >= b1) break;
		const uint32_t y0 = S.segy[s], y1 = S.segy[s + 1];
		const uint32_t adx = x1 - x0;
		const bool up = y1 >= y0;
		const uint32_t ady = up ? y1 - y0 : y0 - y1;
		const float rinv = 1.0f / (float) adx;
		const uint32_t lo = max(x0, b0), hi = min(x1, b1);
		for(uint32_t x = lo + lane; x < hi; x += 32) sink(x, floor1_line_at(x0, y0, adx, ady, up, rinv, x));
	}
	// flat tail from the last flagged post (hpp:583-584)
	const uint32_t xl = S.segx[nseg - 1], yl = S.segy[nseg - 1];
	for(uint32_t x = max(xl, b0) + lane; x < b1; x += 32) sink(x, yl);
}

// ===============================================================================================================
// Fast path used by the fused kernel.
// ===============================================================================================================

// floor(e / d) for e < 2^23 via one float multiply by rinv ~= 1/d and a +/-1 fix-up (exact: e is exactly
// representable, the product is off by far less than 1); larger e (malformed streams only) divide exactly.
__device__ __forceinline__ uint32_t div_floor_u(uint32_t e, uint32_t d, float rinv) {
	if(e >= (1u << 23)) return e / d;
	int q = (int) ((float) e * rinv);
	const int r = (int) e - q * (int) d;
	if(r < 0) --q; else if(r >= (int) d) ++q;
	return (uint32_t) q;
}

// ===============================================================================================================
// v3 curve representation: packed segment records + one segment index per 4-bin cell, evaluated with an exact
// 32.32 fixed-point slope (no division per bin).
//
//   y(x) = y0 +/- floor(k * |dy| / dx),  k = x - x0 < dx < 2^16
//   m = ceil(2^32 * |dy| / dx)  =>  floor(k*m / 2^32) == floor(k*|dy|/dx): the excess k*eps/2^32 (< dx/2^32 < 1/dx)
//   can never carry the fractional part (<= 1 - 1/dx) over the next integer.
// ===============================================================================================================
struct __align__(16) SegRec {
	uint32_t x01;     // x0 | x1 << 16   (x1 = 0xFFFF for the flat tail)
	uint32_t y0s;     // y0 | (descending ? 1u << 31 : 0)
	uint32_t m_lo;    // 32.32 slope magnitude
	uint32_t m_hi;
};

// Per-curve block in shared memory: header (16 B) | SegRec[cap] | uint8 idx[cells], cells = n/8 (4-bin cells over n/2 bins)
struct CurveV3 {
	uint32_t* hdr;      // [0] = mode: 0 curve, 1 multiply by 1.0 (channel untouched), 2 multiply by 0.0 (hpp:1159,1247)
	SegRec*   rec;
	uint8_t*  idx;
	static __host__ __device__ constexpr uint32_t bytes(uint32_t cap, uint32_t cells) { return 16u + cap * 16u + ((cells + 15u) & ~15u); }
	__device__ __forceinline__ void bind(unsigned char* base, uint32_t cap) {
		hdr = reinterpret_cast<uint32_t*>(base);
		rec = reinterpret_cast<SegRec*>(base + 16);
		idx = base + 16 + cap * 16;
	}
};

// Builds the records + cell index from the flagged posts (sx/sy = ascending-x compacted posts held one per lane for
// nseg <= 32, or in shared arrays for larger floors). Lanes cooperate; cells is a multiple of 4.
__device__ __forceinline__ void curve_build_warp(const CurveV3& Cv, uint32_t nseg, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1,
                                                 bool have, uint32_t base_s, uint32_t cells, int lane, bool zero_cells) {
	// this lane's segment s = base_s + lane: from (x0,y0) to (x1,y1); the last one is the flat tail
	if(zero_cells) for(uint32_t w = lane; w < cells / 4; w += 32) reinterpret_cast<uint32_t*>(Cv.idx)[w] = 0;
	__syncwarp();
	if(have) {
		const uint32_t s = base_s + lane;
		const bool last = (s + 1 == nseg);
		SegRec r;
		if(last) { r.x01 = x0 | (0xFFFFu << 16); r.y0s = y0; r.m_lo = 0; r.m_hi = 0; }
		else {
			const bool down = y1 < y0;
			const uint32_t ady = down ? y0 - y1 : y1 - y0, adx = x1 - x0;
			// any m in [v, v+1), v = 2^32*ady/adx, is exact (see header); the rounded-up double quotient (<= 2^-12 above v)
			// plus 0 or the ceil conversion stays inside that interval
			const uint64_t m = __double2ull_ru(__ddiv_ru((double) ady * 4294967296.0, (double) adx));
			r.x01 = x0 | (x1 << 16); r.y0s = y0 | (down ? 0x80000000u : 0u);
			r.m_lo = (uint32_t) m; r.m_hi = (uint32_t) (m >> 32);
		}
		Cv.rec[s] = r;
		// first cell whose first bin is >= x0; write only if no later segment claims the same cell
		const uint32_t cell = (x0 + 3) >> 2;
		const uint32_t ncell = last ? 0xFFFFFFFFu : ((x1 + 3) >> 2);
		if(cell < cells && ncell != cell) Cv.idx[cell] = (uint8_t) s;
	}
}

// inclusive max-scan of the cell index (each cell ends up holding the segment that contains its first bin).
// Four cells per 32-bit word, byte-wise SIMD max (__vmaxu4); cells % 4 == 0.
__device__ __forceinline__ uint32_t bytes_prefix_max(uint32_t v) {      // little-endian: byte i = max(byte 0..i)
	v = __vmaxu4(v, v << 8);
	v = __vmaxu4(v, v << 16);
	return v;
}
__device__ __forceinline__ void curve_scan_cells_warp(const CurveV3& Cv, uint32_t cells, int lane) {
	__syncwarp();
	uint32_t* w = reinterpret_cast<uint32_t*>(Cv.idx);
	const uint32_t words = cells / 4;
	for(uint32_t base = 0, carry = 0; base < words; base += 64) {     // two words per lane per round
		const uint32_t i0 = base + 2 * lane;
		uint32_t a = (i0 < words) ? w[i0] : 0u, b = (i0 + 1 < words) ? w[i0 + 1] : 0u;
		a = bytes_prefix_max(a);
		b = bytes_prefix_max(b);
		b = __vmaxu4(b, (a >> 24) * 0x01010101u);
		uint32_t tot = b >> 24, incl = tot;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
			if(lane >= o) incl = max(incl, v);
		}
		uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
		if(lane == 0) before = 0;
		before = max(before, carry);
		const uint32_t bc = before * 0x01010101u;
		if(i0 < words) w[i0] = __vmaxu4(a, bc);
		if(i0 + 1 < words) w[i0 + 1] = __vmaxu4(b, bc);
		carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
	}
	__syncwarp();
}

// Unwrap (hpp:521-559) + build for floors with <= 32 posts. All 32 lanes call. Returns POV_PKT_* bits (warp-uniform),
// including the hpp:587 range check over the n bins the reference renders.
static __device__ __noinline__ uint32_t floor1_curve_warp32(const DevFloor* __restrict__ F, const uint16_t* __restrict__ ys,
                                                        const CurveV3& Cv, uint32_t cells, uint32_t n, int lane) {
	const int posts = F->n_posts;
	const uint32_t range = F->range;
	const bool have = lane < posts;
	uint32_t cur = have ? ys[lane] : 0u;
	const uint32_t val = cur;
	const int lvl = have ? F->level[lane] : 0;
	const int lo = have ? F->lo[lane] : 0, hi = have ? F->hi[lane] : 0;
	const uint32_t dxn = have ? F->dxn[lane] : 0u, adx = have ? F->adx[lane] : 1u;
	const float rinv = have ? F->rinv[lane] : 1.f;
	uint32_t flags = 0, bad = 0;
	const int levels = F->n_levels;
	for(int lv = 1; lv < levels; ++lv) {
		const uint32_t y0 = __shfl_sync(0xffffffffu, cur, lo), y1 = __shfl_sync(0xffffffffu, cur, hi);
		if(lvl == lv) {
			const bool up = y1 >= y0;
			const uint32_t ady = up ? (y1 - y0) : (y0 - y1);
			const uint32_t off = div_floor_u(ady * dxn, adx, rinv);
			const uint32_t predicted = up ? y0 + off : y0 - off;
			if(predicted > range) bad |= POV_PKT_FLOOR_PREDICTED;
			const uint32_t high_room = range - predicted, low_room = predicted;
			const uint32_t room = min(high_room, low_room) * 2;
			uint32_t fin = predicted;
			if(val != 0) {
				flags |= (1u << lo) | (1u << hi) | (1u << lane);
				if(val >= room) fin = (high_room > low_room) ? val - low_room + predicted : predicted - val + high_room - 1;
				else fin = (val & 1) ? predicted - (val + 1) / 2 : predicted + val / 2;
			}
			cur = fin;
		}
	}
	flags = __reduce_or_sync(0xffffffffu, flags) | 3u;
	// compaction in ascending-x order: lane = sorted position -> lane = segment
	const int si = have ? F->sorted_idx[lane] : 0;
	const uint32_t fy = __shfl_sync(0xffffffffu, cur, si);
	const bool f = have && ((flags >> si) & 1u);
	const uint32_t mask = __ballot_sync(0xffffffffu, f);
	const uint32_t nseg = __popc(mask);
	const uint32_t pos = __popc(mask & ((1u << lane) - 1u));
	const uint32_t myx = f ? (uint32_t) F->xs[si] : 0u;
	const uint32_t myy = f ? min(fy * (uint32_t) F->multiplier, 0xFFFFu) : 0u;
	// lane s takes over flagged post s and s+1 (staged through the record area, which is rewritten right after)
	uint32_t* stage = reinterpret_cast<uint32_t*>(Cv.rec);
	__syncwarp();
	if(f) stage[pos * 4] = myx | (myy << 16);
	__syncwarp();
	const bool seg = (uint32_t) lane < nseg;
	const uint32_t pa = seg ? stage[lane * 4] : 0u, pb = ((uint32_t) lane + 1 < nseg) ? stage[(lane + 1) * 4] : 0u;
	__syncwarp();
	const uint32_t x0 = pa & 0xFFFFu, y0 = pa >> 16, x1 = pb & 0xFFFFu, y1 = pb >> 16;
	// hpp:587 range check: maxima sit at rendered end points (segments are monotone); the segment cut by bin n-1 needs y(n-1)
	bool over = false;
	if(seg) {
		if(x0 < n && y0 >= 256) over = true;
		if((uint32_t) lane + 1 < nseg && x0 < n && x1 > n - 1) {
			const bool down = y1 < y0;
			const uint32_t ady = down ? y0 - y1 : y1 - y0, dx = x1 - x0;
			const uint32_t q = div_floor_u((n - 1 - x0) * ady, dx, 1.0f / (float) dx);
			if((down ? y0 - q : y0 + q) >= 256) over = true;
		}
	}
	if(__any_sync(0xffffffffu, over)) bad |= POV_PKT_FLOOR_RANGE;
	if(lane == 0) Cv.hdr[0] = 0;
	curve_build_warp(Cv, nseg, x0, y0, x1, y1, seg, 0, cells, lane, true);
	curve_scan_cells_warp(Cv, cells, lane);
	return __reduce_or_sync(0xffffffffu, bad);
}

// Two consecutive bins x, x+1 (x even) of the curve as inverse-dB table values.
__device__ __forceinline__ float2 curve_pair(const CurveV3& Cv, uint32_t x, const float* __restrict__ invdb) {
	uint32_t s = Cv.idx[x >> 2];
	uint4 r = *reinterpret_cast<const uint4*>(&Cv.rec[s]);
	while(x >= (r.x >> 16)) r = *reinterpret_cast<const uint4*>(&Cv.rec[++s]);
	float2 out;
	{
		const uint32_t k = x - (r.x & 0xFFFFu);
		const uint32_t q = __umulhi(k, r.z) + k * r.w;
		const uint32_t y0 = r.y & 0xFFFFu;
		const uint32_t y = (r.y >> 31) ? y0 - q : y0 + q;
		out.x = invdb[y & 255];
	}
	const uint32_t xb = x + 1;
	while(xb >= (r.x >> 16)) r = *reinterpret_cast<const uint4*>(&Cv.rec[++s]);
	{
		const uint32_t k = xb - (r.x & 0xFFFFu);
		const uint32_t q = __umulhi(k, r.z) + k * r.w;
		const uint32_t y0 = r.y & 0xFFFFu;
		const uint32_t y = (r.y >> 31) ? y0 - q : y0 + q;
		out.y = invdb[y & 255];
	}
	return out;
}

// Four consecutive bins x..x+3 (x % 4 == 0: the first bin of an index cell) as inverse-dB table values.
static __device__ __noinline__ float4 curve_quad(const CurveV3& Cv, uint32_t x, const float* __restrict__ invdb) {
	uint32_t s = Cv.idx[x >> 2];
	uint4 r = *reinterpret_cast<const uint4*>(&Cv.rec[s]);
	float out[4];
#pragma unroll
	for(int bb = 0; bb < 4; ++bb) {
		const uint32_t xb = x + bb;
		if(bb > 0) while(xb >= (r.x >> 16)) r = *reinterpret_cast<const uint4*>(&Cv.rec[++s]);
		const uint32_t k = xb - (r.x & 0xFFFFu);
		const uint32_t q = __umulhi(k, r.z) + k * r.w;
		const uint32_t y0 = r.y & 0xFFFFu;
		const uint32_t y = (r.y >> 31) ? y0 - q : y0 + q;
		out[bb] = invdb[y & 255];
	}
	return make_float4(out[0], out[1], out[2], out[3]);
}

}  // namespace pov
