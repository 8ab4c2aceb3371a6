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

}  // namespace pov
