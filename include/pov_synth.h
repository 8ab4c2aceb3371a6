/*
 * pov_synth.h — C ABI of the B200-native Vorbis *synthesis stage* (libpov_synth.so).
 *
 * What this boundary replaces in the reference (albertz/ParseOggVorbis, paths relative to its root):
 * the reference has no plugin seam around synthesis — the arithmetic is inlined in
 * VorbisStream::parse_audio (src/ParseOggVorbis.hpp:1128-1274). This header is the seam a maintainer would
 * cut there: everything in parse_audio that does NOT consume the bit reader moves behind these calls, batched
 * over many packets/streams:
 *
 *   floor1 curve synthesis       src/ParseOggVorbis.hpp:521-591, src/Utils.hpp:60-183, src/inverse_db_table.h:13-78
 *   residue VQ application       src/ParseOggVorbis.hpp:685-694, 734-752 (the adds; the Huffman walk stays on the host)
 *   nonzero propagate + coupling src/ParseOggVorbis.hpp:1174-1180, 1213-1241
 *   floor * residue              src/ParseOggVorbis.hpp:1243-1255
 *   inverse MDCT                 src/mdct.h:99-105 (extern "C" mdct_init/mdct_backward/mdct_clear), src/mdct.cpp:433-527
 *   window + overlap-add + emit  src/ParseOggVorbis.hpp:837-886, 1008-1059, 1061-1109
 *
 * Conventions mirror the reference's C API (src/ParseOggVorbis.hpp:1488-1494, src/ParseOggVorbis.cpp:12-41):
 * every call returns int, 0 = ok, non-zero = error; the message is available through pov_last_error() (a
 * per-context buffer instead of the reference's process-wide static one). Nothing throws, nothing aborts on
 * bad data. Only PODs, plain pointers and sizes cross the boundary. All pointers are HOST pointers unless a
 * name says `_dev`. One context per (host thread, GPU); contexts share no mutable state.
 *
 * There is no CPU fallback: every compute entry point fails with POV_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef POV_SYNTH_H
#define POV_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POV_ABI_VERSION      2u
#define POV_MAX_CHANNELS     8u    /* reference: uint8_t audio_channels (hpp:107); this build: <= 8 */
#define POV_MAX_POSTS        256u  /* floor1 syntax bound 2 + 31*8 = 250 (hpp:426,437) */
#define POV_MAX_COUPLINGS    256u  /* hpp:783 */
#define POV_MAX_SUBMAPS      16u   /* hpp:781 */
#define POV_MAX_CLASSES      64u   /* hpp:640 */
#define POV_MAX_MODES        64u   /* hpp:950 */
#define POV_NO_BOOK          255u  /* hpp:658 book_t(-1) */

/* ---- error codes (return values) ------------------------------------------------------------------------- */
enum {
	POV_OK            = 0,
	POV_ERR_ARG       = 1,  /* malformed descriptor / bad argument (host-side validation) */
	POV_ERR_CUDA      = 2,  /* CUDA runtime error or no usable device */
	POV_ERR_STREAM    = 3,  /* a reference CHECK(...) failed while parsing/decoding a stream */
	POV_ERR_UNSUPPORTED = 4 /* valid per the reference but outside this build's limits (e.g. > 8 channels) */
};

/* per-packet device status word (pov_batch_status): which of the reference's fatal per-packet CHECKs fired */
enum {
	POV_PKT_OK               = 0,
	POV_PKT_FLOOR_PREDICTED  = 1, /* hpp:536  CHECK(predicted <= range) */
	POV_PKT_FLOOR_RANGE      = 2, /* hpp:587  CHECK(floor[i] < 256) */
	POV_PKT_VQ_ENTRY         = 4  /* hpp:739,748 decodeVector: entry >= num_entries or lookup_type 0 */
};

/* ---- stream setup (the tables VorbisStreamSetup::parse produces, hpp:889-963) ---------------------------- */
typedef struct pov_codebook {
	uint32_t dim;          /* hpp:121 dimensions_ */
	uint32_t n_entries;    /* hpp:122 num_entries_ */
	uint32_t lookup_type;  /* hpp:142; 0 = scalar-only book, no VQ table */
	uint32_t reserved;
	const float* vq;       /* hpp:149 lookup_table_ [n_entries * dim]; NULL iff lookup_type == 0 */
	const uint8_t* lengths;/* hpp:126 Entry::len_ per entry, 0 = unused entry (hpp:270-272); only read for POV_INPUT_PACKETS
	                          batches (device-side entropy decode), may be NULL otherwise */
} pov_codebook;

/* How a floor1's coded Y list is laid out in the packet (hpp:425-447, read by hpp:498-517): needed only for
 * POV_INPUT_PACKETS batches, where the device walks the packet bits itself. */
typedef struct pov_floor1_syntax {
	uint8_t n_partitions;             /* hpp:426 (<= 31) */
	uint8_t n_classes;                /* hpp:429 maximum class + 1 (<= 16) */
	uint8_t reserved[2];
	uint8_t partition_class[32];      /* hpp:428 */
	uint8_t class_dim[16];            /* hpp:433, 1..8 */
	uint8_t class_subclass_bits[16];  /* hpp:434, 0..3 */
	uint8_t class_masterbook[16];     /* hpp:436 (only if subclass_bits > 0) */
	int16_t class_books[16][8];       /* hpp:441 subclass book number, -1 = none (the Y is 0) */
} pov_floor1_syntax;

typedef struct pov_floor1 {
	uint16_t n_posts;      /* == xs.size(), >= 2 */
	uint8_t  multiplier;   /* 1..4 (hpp:446) */
	uint8_t  reserved;
	uint16_t xs[POV_MAX_POSTS]; /* hpp:448-456, bitstream order; xs[0]=0, xs[1]=1<<rangebits; must be distinct */
} pov_floor1;

typedef struct pov_residue {
	uint32_t type;            /* 0,1,2 (hpp:633) */
	uint32_t begin, end;      /* hpp:636-637 */
	uint32_t partition_size;  /* hpp:639 */
	uint32_t n_class;         /* hpp:640 num_classifications */
	uint32_t classbook;       /* hpp:641 (host entropy decode only; ignored by the device) */
	uint8_t  books[POV_MAX_CLASSES * 8]; /* hpp:652-660 [class*8+pass], POV_NO_BOOK = none */
} pov_residue;

typedef struct pov_mapping {
	uint32_t n_submaps;       /* hpp:779-781 */
	uint32_t n_couplings;     /* hpp:783 */
	uint8_t  mux[POV_MAX_CHANNELS];          /* hpp:795-801 channel -> submap */
	uint8_t  submap_floor[POV_MAX_SUBMAPS];  /* hpp:806 */
	uint8_t  submap_residue[POV_MAX_SUBMAPS];/* hpp:808 */
	uint8_t  coupling_mag[POV_MAX_COUPLINGS];/* hpp:786 */
	uint8_t  coupling_ang[POV_MAX_COUPLINGS];/* hpp:787 */
} pov_mapping;

typedef struct pov_mode {
	uint8_t blockflag;        /* hpp:826: 1 = long window (blocksize1) */
	uint8_t mapping;          /* hpp:831 */
} pov_mode;

typedef struct pov_setup {
	uint32_t abi_version;     /* POV_ABI_VERSION */
	uint32_t channels;        /* hpp:107 */
	uint32_t sample_rate;     /* hpp:108 */
	uint32_t blocksize[2];    /* hpp:114-115; powers of two 64..8192, [0] <= [1] (hpp:1295-1298) */
	uint32_t n_codebooks;  const pov_codebook* codebooks;
	uint32_t n_floors;     const pov_floor1*   floors;
	uint32_t n_residues;   const pov_residue*  residues;
	uint32_t n_mappings;   const pov_mapping*  mappings;
	uint32_t n_modes;      const pov_mode*     modes;
	const pov_floor1_syntax* floor_syntax; /* [n_floors], or NULL (then POV_INPUT_PACKETS batches are refused for this setup) */
} pov_setup;

/* ---- batch description ----------------------------------------------------------------------------------- */
/* One logical stream inside a batch. Its packets are packets[first_packet .. first_packet+n_packets), in
 * decode order. A stream's first packet only primes the overlap (hpp:1021); to continue a stream across
 * batches re-submit the previous batch's last packet first with emit_frames = 0. */
typedef struct pov_stream {
	uint32_t setup_id;        /* from pov_setup_register */
	uint32_t first_packet;
	uint32_t n_packets;
	uint32_t reserved;
	uint64_t pcm_frames;      /* sum of emit_frames of its packets = frames this stream yields in this batch */
	uint64_t pcm_base;        /* float index of channel 0 / frame 0 in the PCM arena. Planar: channel c, frame f
	                             lives at pcm_base + c*pcm_frames + f. Interleaved: pcm_base + f*channels + c. */
} pov_stream;

typedef struct pov_packet {
	uint32_t stream;          /* index into the batch's stream table */
	uint8_t  mode;            /* mode number (hpp:1146) -> blockflag + mapping */
	uint8_t  window_flags;    /* bit0 = prev_window_flag, bit1 = next_window_flag; long blocks only (hpp:1149-1153) */
	uint16_t floor_used;      /* bit c: channel c decoded a floor curve (hpp:478-482), BEFORE nonzero propagate */
	uint32_t emit_frames;     /* frames released by this packet: 0 for a stream's first packet, else
	                             prev/4 + cur/4 (hpp:1026), shortened at a page end by the granule rule (hpp:1028-1033) */
	uint32_t packet_bytes;    /* POV_INPUT_PACKETS: length of the raw audio packet; otherwise unused (0) */
	uint64_t pcm_off;         /* frame index inside the stream's PCM where this packet's chunk starts */
	uint64_t ys_off;          /* index (uint16 units) into the Y arena: for every channel WITH floor_used set, in
	                             channel order, that channel's coded Y list "floor1 ys" (hpp:498-518), n_posts each.
	                             POV_INPUT_PACKETS: ignored (floor_used and the Y lists are decoded on the device) */
	uint64_t spec_off;        /* POV_INPUT_DENSE:   float index into the spectra arena of this packet's
	                                                 "after_residue" vectors, [channels][blocksize/2] (hpp:1211)
	                             POV_INPUT_ENTRIES: byte offset into the residue payload arena (layout below)
	                             POV_INPUT_PACKETS: byte offset (multiple of 4) of the raw audio packet in the payload arena;
	                                                the bytes after it up to the next multiple of 4 must be zero */
} pov_packet;

/* Residue payload of ONE packet in POV_INPUT_ENTRIES mode (everything little-endian, 4-byte aligned):
 *   for each submap s of the packet's mapping, in order (hpp:1184):
 *       uint32 n_entries
 *       uint8  cls[nch_s][parts_s]       classification of every partition (hpp:719), padded to 4 bytes.
 *                                        nch_s = channels muxed to s (1 for residue type 2, hpp:689),
 *                                        parts_s = (min(end,L)-min(begin,L)) / partition_size, L = decode length
 *       uintE  entries[n_entries]        VQ entry numbers in DECODE ORDER (pass, partition, channel, vector:
 *                                        hpp:711-757), padded to 4 bytes. E = 16 bits if every codebook of the
 *                                        setup has <= 65536 entries, else 32 (pov_setup_entry_bits()).
 */
/* POV_INPUT_PACKETS: the payload arena holds the raw Vorbis audio packets (what hpp:1375-1381 hands to parse_audio).
 * mode, window_flags, emit_frames and pcm_off come from the host (they need the first bits of the packet and the page
 * granule only); the device reads everything else from the bits: floor flags and coded Ys (hpp:478-518), residue
 * classifications and VQ entry numbers (hpp:696-760) — the entropy decode of SURVEY.md §8(f)-2. Reading past the end of a
 * packet yields zero bits, as in the reference (Utils.hpp:389-392). The setup must carry codebook lengths and floor syntax. */
enum { POV_INPUT_DENSE = 0, POV_INPUT_ENTRIES = 1, POV_INPUT_PACKETS = 2 };
enum { POV_PCM_PLANAR = 0, POV_PCM_INTERLEAVED = 1 };

typedef struct pov_batch {
	uint32_t input_kind;      /* POV_INPUT_DENSE | POV_INPUT_ENTRIES */
	uint32_t pcm_layout;      /* POV_PCM_PLANAR (matches gotPcmData, hpp:971) | POV_PCM_INTERLEAVED */
	uint32_t n_streams;  const pov_stream* streams;
	uint32_t n_packets;  const pov_packet* packets;
	const uint16_t* ys;       uint64_t n_ys;        /* Y arena */
	const void*     payload;  uint64_t payload_bytes; /* dense spectra (float) or residue payload */
	uint64_t pcm_floats;      /* size of the PCM arena this batch writes (floats) */
} pov_batch;

/* debug stages, named after the reference's dump entries (hpp:518,560,561,585,1171,1211,1254,1265,1051) */
enum {
	POV_STAGE_FINAL_YS = 0,    /* uint32 [posts]      per used channel-packet     "floor1 final_ys"   */
	POV_STAGE_STEP2_FLAG = 1,  /* uint8  [posts]                                   "floor1 step2_flag" */
	POV_STAGE_FLOOR = 2,       /* uint32 [n]                                       "floor1 floor"      */
	POV_STAGE_FLOOR_OUTPUTS = 3,/* float [n]                                       "floor_outputs"     */
	POV_STAGE_AFTER_RESIDUE = 4,/* float [n/2]        per channel-packet           "after_residue"     */
	POV_STAGE_AFTER_ENVELOPE = 5,/* float [n/2]                                    "after_envelope"    */
	POV_STAGE_PCM_AFTER_MDCT = 6 /* float [n]                                      "pcm_after_mdct"    */
};

/* ---- context --------------------------------------------------------------------------------------------- */
typedef struct pov_ctx pov_ctx;

uint32_t    pov_abi_version(void);
/* Create a context on CUDA device `device`. Fails (POV_ERR_CUDA) when no sm_100 GPU is usable. */
int         pov_ctx_create(int device, pov_ctx** out, const char** error_out);
void        pov_ctx_destroy(pov_ctx* ctx);
const char* pov_last_error(const pov_ctx* ctx);
/* The CUDA stream (cudaStream_t as void*) all of this context's work is issued on. */
void*       pov_ctx_stream(pov_ctx* ctx);
/* Number of kernel launches issued by this context so far (for bench accounting). */
uint64_t    pov_ctx_launch_count(const pov_ctx* ctx);
/* Whole-file and corpus decode hand the audio packets to the device as they are (entropy decode in k_packet_decode) when
 * `on` is non-zero — the default; environment POV_DEVICE_ENTROPY=0 turns it off — and walk them on the host otherwise. */
void        pov_ctx_set_device_entropy(pov_ctx* ctx, int on);
/* Whole-file and corpus decode accept packets spanning pages when `on` is non-zero (default off = the reference's hpp:89;
 * environment POV_ALLOW_SPANNING=1 turns it on). */
void        pov_ctx_set_page_spanning(pov_ctx* ctx, int on);
/* Bytes this context has copied host -> device and device -> host so far (descriptor/payload arenas, PCM, status words). */
void        pov_ctx_io_bytes(const pov_ctx* ctx, uint64_t* h2d_out, uint64_t* d2h_out);

/* The 256-entry floor1_inverse_dB_table this library uses (reference: src/inverse_db_table.h:13-78). Host only. */
void        pov_inverse_db_table(float out[256]);
/* The window of a block (hpp:837-862: VorbisModeNumber::precalc / getWindow) as the kernels' slope tables define it:
 * n = blocksize[blockflag] floats. Host-only (no context, no device): the overlap-add stage multiplies by exactly
 * these values. Returns POV_ERR_ARG for block sizes outside 64..8192 / not powers of two / bs0 > bs1 / n mismatch. */
int         pov_window(uint32_t blocksize0, uint32_t blocksize1, int blockflag, int prev, int next, float* out, uint32_t n);

/* Validate + upload one stream setup; derived tables (neighbours, sort order, windows, twiddles) are built here. */
int         pov_setup_register(pov_ctx* ctx, const pov_setup* setup, uint32_t* setup_id_out);
int         pov_setup_entry_bits(const pov_ctx* ctx, uint32_t setup_id); /* 16 or 32; <0 on bad id */
/* Copy the window table the library generated for (setup, blockflag, prev, next) — hpp:837-886 — n floats. */
int         pov_setup_get_window(const pov_ctx* ctx, uint32_t setup_id, int blockflag, int prev, int next,
                                 float* out, uint32_t n);

/* ---- batches --------------------------------------------------------------------------------------------- */
typedef struct pov_batch_handle pov_batch_handle;

/* Validate a batch, build the device work lists and copy descriptors + arenas host->device (async on the
 * context stream; the host buffers must stay valid until pov_batch_sync or the next synchronising call). */
int  pov_batch_upload(pov_ctx* ctx, const pov_batch* batch, pov_batch_handle** out);
/* Launch the production path: [residue apply] -> fused floor1/coupling/dot/IMDCT/window/OLA -> PCM arena (device). */
int  pov_batch_run(pov_ctx* ctx, pov_batch_handle* h);
/* Name of the kernel pov_batch_run launches for this batch: "k_warp_synth" (persistent warp-autonomous kernel: block
 * sizes 256/512/1024/2048, <= 32 posts per floor), "k_fused_synth" (one CTA per run, any block sizes) or "staged".
 * Static string. */
const char* pov_batch_kernel_name(const pov_ctx* ctx, const pov_batch_handle* h);
/* Launch the staged path instead: one kernel per reference stage, every intermediate materialised in HBM so
 * that pov_batch_fetch_stage can return it. Produces the same PCM. */
int  pov_batch_run_staged(pov_ctx* ctx, pov_batch_handle* h);
/* Device->host copy of the PCM arena (pcm_floats floats; async unless `sync`). */
int  pov_batch_fetch_pcm(pov_ctx* ctx, pov_batch_handle* h, float* out, uint64_t n_floats, int sync);
/* Device pointer of the PCM arena (for callers that consume PCM on the GPU). */
void* pov_batch_pcm_dev(pov_batch_handle* h);
/* One debug stage of one (packet, channel) after pov_batch_run_staged; `out_bytes` must match the stage size. */
int  pov_batch_fetch_stage(pov_ctx* ctx, pov_batch_handle* h, uint32_t packet, uint32_t channel, int stage,
                           void* out, uint64_t out_bytes);
/* Per-packet status words (POV_PKT_*), n_packets entries; synchronises. Returns POV_ERR_STREAM if any is non-zero. */
int  pov_batch_status(pov_ctx* ctx, pov_batch_handle* h, uint32_t* out_status, uint32_t n);
int  pov_batch_sync(pov_ctx* ctx, pov_batch_handle* h);
void pov_batch_free(pov_ctx* ctx, pov_batch_handle* h);

/* ---- drop-in for the reference's only C symbols on the path: mdct_backward (src/mdct.h:105) ---------------- */
/* `count` independent inverse MDCTs of size n (R^(n/2) -> R^n, no scaling), host buffers in/out. */
/* Parity hook of the production kernel: its own floor1 step-1 result (what hpp:521-559 computes), [n_packets][max channels][72]
 * bytes: 64 final Ys in ascending-x order, clamped to 255 | 64-bit step-2 mask in the same order. Warp-kernel batches only. */
int  pov_batch_fetch_fast_floor(pov_ctx* ctx, pov_batch_handle* h, uint8_t* out, uint64_t out_bytes);

/* Feature matrices straight from the device (the reference's downstream use: returnn_import.py:74-115 builds them in
 * Python from a debug dump, demo_live_extract.py:262-505): `kind` as below with the readers' default arguments, for one
 * stream of the batch; (rows, output_dim) floats, row-major, only the matrix crosses PCIe. Runs the staged kernels if they
 * have not run yet. out == NULL: only *rows_out is set (size query).
 *   FLOOR_FINAL_YS           get_features_from_raw_bytes(kind="floor_final_ys")            one row per decoded floor curve
 *   FLOOR_FINAL_YS_RENDERED  kind="floor_final_ys_rendered"
 *   RESIDUE_YS               kind="residue_ys"               rows of the packets whose (last channel's) floor has most posts
 *   RESIDUE_YS_WITH_FLOOR    kind="residue_ys_with_floor"    (exp() is the device's: within 1e-6 relative of numpy's) */
#define POV_ALL_STREAMS 0xFFFFFFFFu   /* `stream`: the rows of every stream of the batch, one after the other (one setup) */
enum { POV_FEAT_FLOOR_FINAL_YS = 0, POV_FEAT_FLOOR_FINAL_YS_RENDERED = 1, POV_FEAT_RESIDUE_YS = 2, POV_FEAT_RESIDUE_YS_WITH_FLOOR = 3 };
int  pov_batch_features(pov_ctx* ctx, pov_batch_handle* h, uint32_t stream, int kind, uint32_t output_dim,
                        float* out, uint64_t rows_cap, uint64_t* rows_out);

int  pov_mdct_backward_batch(pov_ctx* ctx, uint32_t n, uint64_t count, const float* in, float* out);

/* ---- whole-stream decode through the host front-end (mirrors ogg_vorbis_full_read_from_memory, hpp:1493) - */
typedef struct pov_decoded {
	uint32_t channels;
	uint32_t sample_rate;
	uint64_t frames;          /* frames per channel */
	uint32_t audio_packets;
	uint32_t reserved;
	float*   pcm;             /* planar [channels][frames]; owned by the library, free with pov_decoded_free */
} pov_decoded;

/* Decode one Ogg/Vorbis file held in memory. `debug_out` (may be NULL) names a file that receives the reference's
 * "ParseOggVorbis-header-v1" dump (src/Callbacks.cpp:146-185), produced from the staged device path. */
int  pov_ogg_vorbis_decode_memory(pov_ctx* ctx, const uint8_t* data, size_t len, const char* debug_out,
                                  pov_decoded* out);
void pov_decoded_free(pov_decoded* d);

/* Decode a corpus of independent files with `host_threads` front-end threads (0 = all cores but one, which is left to
 * the calling thread) feeding this context's GPU. The context keeps the sibling stream, device arenas and pinned
 * staging buffers of the call for the next one; POV_CORPUS_TIMING=1 prints where the calling thread's time went.
 * Returns per-file frame counts (frames_out[n_files], may be NULL) and the total number of PCM values. */
int  pov_decode_corpus(pov_ctx* ctx, uint32_t n_files, const uint8_t* const* data, const size_t* len,
                       uint32_t host_threads, uint64_t* frames_out, uint64_t* total_values_out,
                       double* checksum_out);

/* The same decode with an output edge: `sink` receives the PCM of every logical stream, in file order, from the calling
 * thread, as the chunk that holds it retires — the batch form of ParseCallbacks::gotPcmData (hpp:966-973, 1047-1053).
 * `planar` is [channels][frames] (channel c at planar + c * frames) and is valid during the call only. A non-zero return
 * stops the decode, which then fails with POV_ERR_STREAM like the reference's CHECK(callbacks.gotPcmData(...)). */
typedef int (*pov_pcm_sink)(uint32_t file_index, uint32_t channels, uint64_t frames, const float* planar, void* user);
int  pov_decode_corpus_pcm(pov_ctx* ctx, uint32_t n_files, const uint8_t* const* data, const size_t* len,
                           uint32_t host_threads, pov_pcm_sink sink, void* user,
                           uint64_t* frames_out, uint64_t* total_values_out, double* checksum_out);

/* Host-only front-end (no GPU needed): parse a whole Ogg/Vorbis file into descriptor batches, one per logical stream
 * (Ogg framing hpp:51-102,1433-1484; headers hpp:1283-1373; per-packet entropy decode hpp:498-517, 711-757). The
 * setup/batch returned by pov_parsed_get point into memory owned by the handle (POV_INPUT_ENTRIES, setup_id 0). */
typedef struct pov_parsed pov_parsed;
int      pov_ogg_parse_memory(const uint8_t* data, size_t len, pov_parsed** out, const char** error_out);
/* flags: POV_PARSE_RAW_PACKETS — do not entropy-decode the audio packets; the batches are POV_INPUT_PACKETS (a stream whose
 * setup cannot be walked on the device — floor0, a submap without channels — still comes back as POV_INPUT_ENTRIES). */
#define POV_PARSE_RAW_PACKETS 1u
/*        POV_PARSE_ALLOW_SPANNING — accept packets that continue on the next page of their stream (RFC 3533). The reference
 * refuses such files (hpp:89: "we don't support packets spanning pages"), which stays the default for parity. */
#define POV_PARSE_ALLOW_SPANNING 2u
int      pov_ogg_parse_memory_ex(const uint8_t* data, size_t len, uint32_t flags, pov_parsed** out, const char** error_out);
uint32_t pov_parsed_stream_count(const pov_parsed* p);
int      pov_parsed_get(const pov_parsed* p, uint32_t stream, pov_setup* setup_out, pov_batch* batch_out);
void     pov_parsed_free(pov_parsed* p);

/* Same signature/semantics as the reference's entry point (hpp:1493), on device 0; error string is per-thread. */
int  pov_ogg_vorbis_full_read_from_memory(const char* data, size_t data_len, const char** error_out);

#ifdef __cplusplus
}
#endif
#endif /* POV_SYNTH_H */
