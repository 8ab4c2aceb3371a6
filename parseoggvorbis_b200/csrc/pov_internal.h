// Internal structures shared by the host side and the kernels of libpov_synth.so.
// Device-side tables are flat PODs built once per setup by pov_setup_register (host_tables.cpp).
#ifndef POV_INTERNAL_H
#define POV_INTERNAL_H

#include <stdint.h>
#include "../../include/pov_synth.h"

#define POV_MAX_CODEBOOKS 256

// floor1 tables derived from the setup's X list (reference: hpp:458-469 sort, Utils.hpp:60-118 neighbours).
struct DevFloor {
	uint16_t n_posts;
	uint8_t  multiplier;
	uint8_t  n_levels;                 // depth of the neighbour dependency DAG (+1)
	uint32_t range;                    // hpp:486-492
	uint16_t xs[POV_MAX_POSTS];        // bitstream order
	uint8_t  lo[POV_MAX_POSTS];        // low_neighbor index (posts >= 2)
	uint8_t  hi[POV_MAX_POSTS];        // high_neighbor index
	uint8_t  level[POV_MAX_POSTS];     // 0 for posts 0,1; else 1+max(level[lo],level[hi])
	uint8_t  sorted_idx[POV_MAX_POSTS];// ascending-x order -> post index
	// prediction constants of post i (>= 2): predicted = y_lo +/- floor(|y_hi-y_lo| * dxn / adx)   (Utils.hpp:122-137)
	uint16_t dxn[POV_MAX_POSTS];       // xs[i] - xs[lo]
	uint16_t adx[POV_MAX_POSTS];       // xs[hi] - xs[lo]
	float    rinv[POV_MAX_POSTS];      // 1.0f / adx
};

struct DevMapping {
	uint32_t n_submaps;
	uint32_t n_couplings;
	uint8_t  mux[POV_MAX_CHANNELS];
	uint8_t  floor_of_ch[POV_MAX_CHANNELS];   // submap_floor[mux[c]]
	uint8_t  submap_residue[POV_MAX_SUBMAPS];
	uint8_t  coupling_mag[POV_MAX_COUPLINGS];
	uint8_t  coupling_ang[POV_MAX_COUPLINGS];
};

struct DevResidue {
	uint32_t type, begin, end, partition_size, n_class;
	uint8_t  books[POV_MAX_CLASSES * 8];
};

struct DevCodebook {
	uint32_t dim, n_entries, lookup_type, pad;
	const float* vq;                   // device pointer
};

// ---- device-side entropy decode (POV_INPUT_PACKETS, kernel_entropy.cu) ------------------------------------------------
// Huffman tables of one codebook inside the setup's uint32 arena: a first-level table indexed by the next kHuffLutBits
// stream bits ((entry << 6) | length; 0 = the codeword is longer) and, for those, the codewords sorted by left-aligned value
// (sorted_code[i], then (entry << 6 | length)[i]).
#define POV_HUFF_LUT_BITS 9
struct DevHuffBook {
	uint32_t lut_off;       // uint32 index into the arena
	uint32_t sorted_off;    // n_sorted codes, then n_sorted (entry << 6 | length)
	uint32_t n_sorted;
	uint32_t n_entries;
	uint32_t dim;
	uint32_t lookup_type;
	uint32_t pad[2];
};
struct DevFloorSyntax {         // pov_floor1_syntax + what the Y walk needs from the floor itself
	uint8_t n_partitions, n_classes, ybits, pad;
	uint8_t partition_class[32];
	uint8_t class_dim[16], class_subclass_bits[16], class_masterbook[16];
	int16_t class_books[16][8];
};

struct DevSetup {
	uint32_t channels;
	uint32_t blocksize[2];
	uint32_t log2bs[2];
	uint32_t n_floors, n_mappings, n_modes, n_residues, n_codebooks;
	uint32_t entry_bits;               // 16 or 32
	uint8_t  mode_blockflag[POV_MAX_MODES];
	uint8_t  mode_mapping[POV_MAX_MODES];
	const DevFloor*    floors;         // device arrays
	const DevMapping*  mappings;
	const DevResidue*  residues;
	const DevCodebook* codebooks;
	// entropy decode on the device (null when the setup came without codebook lengths / floor syntax)
	const uint32_t*       huff;        // arena
	const DevHuffBook*    hbooks;      // [n_codebooks]
	const DevFloorSyntax* fsyntax;     // [n_floors]
	uint32_t classbook[64];            // per residue (pov_residue::classbook)
	uint32_t mode_bits;                // ilog(n_modes - 1): bits of the mode number (hpp:1146)
	uint32_t pad_e;
	// Rising window slope of length blocksize[k]/2 (hpp:850-853); the falling slope is its mirror image
	// (same argument expression, hpp:857) and everything else of a window is exactly 0 or 1.
	const float*  slope[2];
	// DCT-IV twiddles w[j] = exp(-i*pi*(8j+1)/(8M)), j < M/2, M = blocksize/2  (pre- and post-rotation)
	const float2* rot[2];
	// FFT twiddles W_Q^e = exp(-2*pi*i*e/Q), e < Q, Q = blocksize/4
	const float2* fft[2];
	// The same twiddles laid out per pass for the fused kernel: for every DIF pass (first small-radix pass, then
	// the radix-8 passes with L >= 64) the 8 (or 4 / 1) factors of butterfly j are contiguous (fft_core.cuh PassTables)
	const float2* fftp[2];
	// Compact tables the fused kernel copies into shared memory: for every radix-8 pass with L >= 64 the factors
	// W_L^j, W_L^2j, W_L^3j, W_L^4j of butterfly j (the other three are one multiplication away), fft8_count float2 each
	const float2* fft8[2];
	uint32_t fft8_count[2];
	// rotation helpers: rotc1 = exp(-i*pi/M) (rot[j+1] = rot[j]*rotc1), rotc6 = exp(+i*pi*6/(8M))
	// (rot[Q-1-j] = -i * conj(rot[j]) * rotc6)
	float2 rotc1[2], rotc6[2];
};

// ---- compact per-setup tables of the warp-autonomous kernel (kernel_warp.cu), copied into shared memory once per CTA --
#define POV_FAST_MAX_FLOORS   4
#define POV_FAST_MAX_MAPPINGS 4
#define POV_FAST_MAX_STEPS    8     // coupling steps per mapping
#define POV_FAST_MAX_DEPS     4     // channels one channel's un-coupled value may depend on (itself included)
#define POV_FAST_MAX_POSTS    64    // floor posts on the warp kernel's path (libvorbis' floors have up to 65; the fixtures' 9 / 29)
#define POV_FAST_MAX_X        1024  // largest floor X (segment lengths index the reciprocal table)

struct FastFloor {
	uint32_t n_posts, n_levels, range, multiplier;
	// per lane i (16 bytes, one LDS.128):
	//   [0] lo | hi << 8 | level << 16 | sorted_idx << 24     (post i; sorted_idx: post index of the i-th smallest X)
	//   [1] (xs[i] - xs[lo]) | (xs[hi] - xs[lo]) << 16         (prediction numerator / denominator, Utils.hpp:122-137)
	//   [2] ceil(2^32 / (xs[hi] - xs[lo]))                     (exact division by multiplication, see kernel_warp.cu)
	//   [3] X of the i-th smallest post
	uint32_t post[POV_FAST_MAX_POSTS][4];
	uint32_t xs_sorted[POV_FAST_MAX_POSTS];   // X of the i-th smallest post | its post number << 16, unit stride (conflict-free per-lane reads)
};

// How the warp that owns channel c of a mapping un-couples it (hpp:1213-1241): the channels it has to load
// (ch[0] == c) and the sub-sequence of coupling steps, in application order, that can reach c; indices are local.
struct FastCouple {
	uint8_t nl, nsteps;
	uint8_t ch[POV_FAST_MAX_DEPS];
	uint8_t sm[POV_FAST_MAX_STEPS], sa[POV_FAST_MAX_STEPS];
	uint8_t pad[2];
};

struct FastTables {
	FastFloor  floors[POV_FAST_MAX_FLOORS];
	FastCouple couple[POV_FAST_MAX_MAPPINGS][POV_MAX_CHANNELS];
	uint8_t floor_of_ch[POV_FAST_MAX_MAPPINGS][POV_MAX_CHANNELS];
	uint8_t ncoup[POV_FAST_MAX_MAPPINGS];                           // full coupling list, for the nonzero propagate rule (hpp:1174-1180)
	uint8_t cmag[POV_FAST_MAX_MAPPINGS][POV_FAST_MAX_STEPS], cang[POV_FAST_MAX_MAPPINGS][POV_FAST_MAX_STEPS];
	uint8_t mode_flag[POV_MAX_MODES], mode_map[POV_MAX_MODES];
	uint32_t channels;
	uint32_t short_posts_cap;          // record capacity of a short-block curve (multiple of 4)
	uint32_t long_posts_cap;           // record capacity of a long-block curve (32 or 64)
	uint32_t wide;                     // a reachable floor has more than 32 posts (64-bit step-2 masks, 72-byte Y records)
	uint32_t pad[3];
};
static_assert(sizeof(FastTables) % 16 == 0, "FastTables is moved with 16-byte bulk copies");

// Work item of the fused kernel: a run of consecutive packets of one stream. The first packet of a run that is
// not the first packet of its stream is a halo: it is transformed again only to rebuild the overlap half.
struct DevRun {
	uint32_t first_packet;             // global packet index where processing starts (halo included)
	uint32_t n_packets;                // packets processed, halo included
	uint32_t halo;                     // 1 -> first processed packet emits nothing here
	uint32_t pad;
};

#endif
