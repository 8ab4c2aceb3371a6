// Host-side generation of the tables the kernels consume (built once per setup / per blocksize).
#ifndef POV_HOST_TABLES_H
#define POV_HOST_TABLES_H

#include <stdint.h>
#include <string>
#include <vector>
#include "pov_internal.h"

namespace pov {

// floor1_inverse_dB_table (reference: src/inverse_db_table.h:13-78; Vorbis I spec 10.1)
void make_inverse_db_table(float out[256]);
// Rising window slope of `len` samples (reference: src/ParseOggVorbis.hpp:850-853)
void make_window_slope(uint32_t len, std::vector<float>& out);
// Full window of a block (reference: src/ParseOggVorbis.hpp:837-862), for pov_setup_get_window
void make_window(uint32_t bs0, uint32_t bs1, int blockflag, int prev, int next, std::vector<float>& out);
// DCT-IV rotation w[j] = exp(-i*pi*(8j+1)/(8M)), j < M/2, M = n/2; FFT twiddles W_Q^e, Q = n/4
void make_rotation(uint32_t n, std::vector<float>& out_re_im);
void make_fft_twiddles(uint32_t n, std::vector<float>& out_re_im);
// per-pass packed twiddles (layout: fft_core.cuh PassTables<Q>)
void make_fft_pass_tables(uint32_t n, std::vector<float>& out_re_im);
// compact radix-8 pass tables (layout: fft_core.cuh Tw8Tables<Q>) and the rotation helper constants
void make_fft_r8_tables(uint32_t n, std::vector<float>& out_re_im);
void make_rotation_consts(uint32_t n, float c1[2], float c6[2]);
// Per-lane factor rows of the 512-point register FFT whose lane <-> register exchanges go through tensor memory
// (kernel_warp.cu "fft512_tm"): [32 lanes][kTmTableCols] floats, block size n = 2048 only. Column map in kernel_warp.cu.
constexpr uint32_t kTmTableCols = 144;
void make_tm_lane_tables(uint32_t n, const std::vector<float>& rot_re_im, const std::vector<float>& slope, std::vector<float>& out);
// index maps of that FFT (shared by the table generator and the tests' model): output frequency held by (lane, register)
uint32_t tm_fft_freq_of(uint32_t lane, uint32_t reg);
// floor1 derived tables. Returns false (with msg) when the X list is not usable.
bool make_floor_tables(const pov_floor1& in, DevFloor& out, std::string& msg);

}  // namespace pov
#endif
