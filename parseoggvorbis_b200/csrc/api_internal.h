// Host-side state behind the opaque handles of include/pov_synth.h.
#ifndef POV_API_INTERNAL_H
#define POV_API_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "pov_internal.h"

struct DevBuf {
	void* ptr = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t bytes);
	void release();
};

struct BlockTables {              // per blocksize n: DCT-IV rotation, FFT twiddles, rising slope of n/2 samples
	const float2* d_rot = nullptr;
	const float2* d_fft = nullptr;
	const float2* d_fftp = nullptr;
	const float2* d_fft8 = nullptr;
	uint32_t fft8_count = 0;
	float c1[2] = {1, 0}, c6[2] = {1, 0};
	const float*  d_slope = nullptr;
	std::vector<float> h_slope;
	const float*  d_tm = nullptr;     // n == 2048: lane rows of the tensor-memory FFT (make_tm_lane_tables)
};

struct SetupRec {
	uint32_t channels = 0, sample_rate = 0, blocksize[2] = {0, 0};
	uint32_t n_modes = 0, entry_bits = 16, max_posts = 2, res_smem = 0;
	uint32_t posts_cls[2] = {2, 2};   // largest floor (posts) reachable from short / long modes
	uint32_t table_float2 = 0;        // float2 slots of the fused kernel's shared-memory tables for this setup
	uint8_t mode_blockflag[POV_MAX_MODES] = {0};
	uint8_t mode_mapping[POV_MAX_MODES] = {0};
	std::vector<DevFloor> floors_host;
	std::vector<DevMapping> maps_host;
	std::vector<DevResidue> residues_host;
	std::vector<uint32_t> cb_dim;
	DevSetup dev;
	const DevFloor* d_floors = nullptr;
	const DevMapping* d_mappings = nullptr;
	const DevResidue* d_residues = nullptr;
	const DevCodebook* d_codebooks = nullptr;
	const float* d_vq = nullptr;
	// warp-autonomous kernel (kernel_warp.cu): eligible setups carry their compact tables
	bool fast_ok = false;
	uint32_t fast_short_cap = 4, fast_long_cap = 32;
	bool fast_wide = false;           // a reachable floor has 33..64 posts: shorter runs (72-byte Y records)
	uint32_t fast_max_nl = 1;         // largest channel set a coupling program of this setup needs
	const FastTables* d_fast = nullptr;
	// device entropy decode (POV_INPUT_PACKETS): tables + per-mode arena capacities of one packet
	bool entropy_ok = false;
	const uint32_t* d_huff = nullptr;
	const DevHuffBook* d_hbooks = nullptr;
	const DevFloorSyntax* d_fsyntax = nullptr;
	uint32_t mode_ys_cap[POV_MAX_MODES] = {0};       // uint16 slots: every channel's Y list
	uint32_t mode_ent_cap[POV_MAX_MODES] = {0};      // bytes of the entries payload in the worst case (multiple of 4)
	std::string image;            // canonical bytes, for de-duplication
};

struct pov_ctx {
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	// Asynchronous copies back to the host run on a stream of their own, joined with `stream` by events: an operation that
	// follows a device-to-host copy in stream order is submitted to the copy engine's queue, i.e. behind every copy-out that
	// any other stream submitted before it (measured: the next chunk's copy-in + kernels waited for three foreign copy-outs).
	cudaStream_t out_stream = nullptr;
	cudaEvent_t ev_compute = nullptr, ev_out = nullptr;
	uint64_t launches = 0;
	uint64_t h2d_bytes = 0, d2h_bytes = 0;   // what this context copied between host and device (pov_ctx_io_bytes)
	uint32_t run_len = 0;         // 0 = automatic
	int kernel_choice = 0;        // POV_KERNEL: 0 automatic, 1 force the CTA-per-run fused kernel, 2 require the warp kernel
	bool allow_spanning = false;  // accept packets that span pages (the reference refuses them, hpp:89)
	bool device_entropy = true;   // whole-file / corpus decode: audio packets are entropy-decoded on the device (POV_DEVICE_ENTROPY=0: on the host)
	uint32_t* d_counter = nullptr; // work counter of the persistent kernel
	const float* d_inv_db = nullptr;
	const DevSetup* d_setups = nullptr;
	std::vector<SetupRec> setups;
	std::map<uint32_t, BlockTables> blk_tables;
	std::map<std::string, uint32_t> setup_by_key;   // raw header bytes of a parsed stream -> setup id (front end)
	DevBuf mdct_in, mdct_out;
	void* corpus = nullptr;                 // what pov_decode_corpus keeps between calls (front_end.cpp: sibling context, slots, pinned pool)
	void (*corpus_free)(void*) = nullptr;
	char err[512] = {0};
};

struct WarpGroup { uint32_t setup, first_run, n_runs; };   // runs of one setup: one launch of the persistent warp kernel

struct pov_batch_handle {
	uint32_t n_streams = 0, n_packets = 0, input_kind = 0, pcm_layout = 0;
	uint64_t pcm_floats = 0, stage_floats = 0, dense_floats = 0;
	uint32_t max_channels = 1, max_blocksize = 64, min_blocksize = 64, floor_cap = 4, res_smem = 0;
	uint32_t floor_cap_cls[2] = {4, 4};
	uint32_t table_float2 = 0;
	bool fused_ok = true, staged_ready = false, only_256_2048 = false;
	bool warp_ok = false;         // every stream uses one setup that kernel_warp.cu supports
	uint32_t warp_setup = 0;
	std::vector<WarpGroup> warp_groups;
	std::vector<uint64_t> spec_off, stage_off;
	std::vector<uint64_t> pk_ys_off, pk_ent_off, pk_raw_off;     // POV_INPUT_PACKETS: capacity-based places of a packet's Y lists / entries payload
	uint64_t ys_cap = 0, ent_cap = 0;
	std::vector<uint32_t> pk_n, pk_setup;
	std::vector<uint16_t> pk_used;                   // floor_used as the host saw it (POV_INPUT_PACKETS: decoded on the device instead)
	std::vector<uint8_t> pk_mode;
	std::vector<pov_stream> streams_host;
	DevBuf d_feat_rows, d_feat_floors, d_feat_out;   // pov_batch_features
	std::vector<DevRun> runs;
	DevBuf d_streams, d_packets, d_ys, d_payload, d_spec_off, d_stage_off, d_runs, d_pcm, d_status, d_spectra;
	DevBuf d_entries, d_pk_off;                      // POV_INPUT_PACKETS: entries payload written by k_packet_decode; ys_off | ent_off
	DevBuf st_final_ys, st_flag, st_floor, st_floor_out, st_env, st_mdct;
	// pinned copy of the derived arrays (spec_off | runs): sources of asynchronous copies, so they must not be pageable
	// (a pageable source makes the "async" copy wait for the stream) and must outlive the copy (derived_copied)
	void* h_derived = nullptr; size_t h_derived_cap = 0;
	cudaEvent_t derived_copied = nullptr;
};

struct StageHost {                // whole stage arrays on the host (debug dump writer)
	std::vector<uint32_t> final_ys;
	std::vector<uint8_t> step2;
	std::vector<uint16_t> floor;
	std::vector<float> floor_out, env, mdct, residue;
};

// Device -> host copy on the context's copy-out stream. fork: ordered after everything queued on ctx->stream so far; join:
// work queued on ctx->stream afterwards waits for the copy (and cudaStreamSynchronize(ctx->stream) covers it).
cudaError_t pov_copy_out_async(pov_ctx* ctx, void* dst, const void* src, size_t bytes, bool fork = true, bool join = true);
int pov_fail(pov_ctx* ctx, int code, const char* fmt, ...);
// Entry points are function-try-blocks that end in POV_NOTHROW_END: no exception crosses the C boundary.
int pov_fail_exception(pov_ctx* ctx);
#define POV_NOTHROW_END(ctx) catch(...) { return pov_fail_exception(ctx); }
int pov_batch_fetch_stage_all(pov_ctx* ctx, pov_batch_handle* h, StageHost& out);

#endif
