// Whole-file / corpus decode: host front-end (vorbis_parse.cpp) -> descriptor batch -> C ABI -> CUDA kernels.
// Mirrors the reference's entry points ogg_vorbis_full_read_from_memory (src/ParseOggVorbis.hpp:1493,
// src/ParseOggVorbis.cpp:28-41) and OggReader::full_read_from_memory (hpp:1428): 0 = ok, non-zero = error + message.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "api_internal.h"
#include "debug_dump.h"
#include "kernels.h"
#include "vorbis_parse.h"

using namespace pov;

// ---------------------------------------------------------------------------------------------------------------
// parsed file handle (host only; usable without a GPU, e.g. from worker processes)
// ---------------------------------------------------------------------------------------------------------------
struct pov_parsed {
	std::vector<StreamWork> streams;
	std::vector<std::unique_ptr<SetupAbi>> abi;        // one per stream
	std::vector<pov_stream> stream_rec;                // single-stream batches
};

static bool finish_stream_tables(pov_parsed& P, std::string& why) {
	P.abi.clear();
	P.stream_rec.clear();
	for(StreamWork& st : P.streams) {
		std::unique_ptr<SetupAbi> a(new SetupAbi());
		if(!st.have_setup) {
			memset(&a->s, 0, sizeof a->s);
		} else if(!setup_to_abi(st.setup, *a, why)) {
			return false;
		}
		P.abi.push_back(std::move(a));
		pov_stream rec;
		memset(&rec, 0, sizeof rec);
		rec.setup_id = 0; rec.first_packet = 0; rec.n_packets = (uint32_t) st.packets.size();
		rec.pcm_frames = st.frames; rec.pcm_base = 0;
		P.stream_rec.push_back(rec);
		for(pov_packet& pk : st.packets) pk.stream = 0;
	}
	return true;
}

extern "C" int pov_ogg_parse_memory(const uint8_t* data, size_t len, pov_parsed** out, const char** error_out) {
	static thread_local char errbuf[512];
	if(!out || (!data && len)) return POV_ERR_ARG;
	*out = nullptr;
	std::unique_ptr<pov_parsed> P(new pov_parsed());
	ParseError err;
	if(!parse_ogg_file(data, len, P->streams, err)) {
		snprintf(errbuf, sizeof errbuf, "check failed: %s", err.msg.c_str());
		if(error_out) *error_out = errbuf;
		return POV_ERR_STREAM;
	}
	std::string why;
	if(!finish_stream_tables(*P, why)) {
		snprintf(errbuf, sizeof errbuf, "unsupported stream: %s", why.c_str());
		if(error_out) *error_out = errbuf;
		return POV_ERR_UNSUPPORTED;
	}
	*out = P.release();
	return POV_OK;
}

extern "C" uint32_t pov_parsed_stream_count(const pov_parsed* p) { return p ? (uint32_t) p->streams.size() : 0; }

extern "C" int pov_parsed_get(const pov_parsed* p, uint32_t stream, pov_setup* setup_out, pov_batch* batch_out) {
	if(!p || stream >= p->streams.size()) return POV_ERR_ARG;
	const StreamWork& st = p->streams[stream];
	if(setup_out) *setup_out = p->abi[stream]->s;
	if(batch_out) {
		memset(batch_out, 0, sizeof *batch_out);
		batch_out->input_kind = POV_INPUT_ENTRIES;
		batch_out->pcm_layout = POV_PCM_PLANAR;
		batch_out->n_streams = 1; batch_out->streams = &p->stream_rec[stream];
		batch_out->n_packets = (uint32_t) st.packets.size(); batch_out->packets = st.packets.data();
		batch_out->ys = st.ys.data(); batch_out->n_ys = st.ys.size();
		batch_out->payload = st.payload.data(); batch_out->payload_bytes = st.payload.size();
		batch_out->pcm_floats = st.frames * st.setup.channels;
	}
	return POV_OK;
}

extern "C" void pov_parsed_free(pov_parsed* p) { delete p; }

// ---------------------------------------------------------------------------------------------------------------
// one batch out of several parsed streams
// ---------------------------------------------------------------------------------------------------------------
struct HostBatch {
	std::vector<pov_stream> streams;
	std::vector<pov_packet> packets;
	std::vector<uint16_t> ys;
	std::vector<uint8_t> payload;
	uint64_t pcm_floats = 0;
	pov_batch view() const {
		pov_batch b;
		memset(&b, 0, sizeof b);
		b.input_kind = POV_INPUT_ENTRIES; b.pcm_layout = POV_PCM_PLANAR;
		b.n_streams = (uint32_t) streams.size(); b.streams = streams.data();
		b.n_packets = (uint32_t) packets.size(); b.packets = packets.data();
		b.ys = ys.data(); b.n_ys = ys.size();
		b.payload = payload.data(); b.payload_bytes = payload.size();
		b.pcm_floats = pcm_floats;
		return b;
	}
	void append(const StreamWork& st, uint32_t setup_id) {
		pov_stream rec;
		memset(&rec, 0, sizeof rec);
		rec.setup_id = setup_id; rec.first_packet = (uint32_t) packets.size(); rec.n_packets = (uint32_t) st.packets.size();
		rec.pcm_frames = st.frames; rec.pcm_base = pcm_floats;
		const uint32_t si = (uint32_t) streams.size();
		const uint64_t ybase = ys.size(), pbase = payload.size();
		for(const pov_packet& src : st.packets) {
			pov_packet pk = src;
			pk.stream = si; pk.ys_off += ybase; pk.spec_off += pbase;
			packets.push_back(pk);
		}
		ys.insert(ys.end(), st.ys.begin(), st.ys.end());
		payload.insert(payload.end(), st.payload.begin(), st.payload.end());
		pcm_floats += st.frames * st.setup.channels;
		streams.push_back(rec);
	}
};

static int register_stream_setup_slow(pov_ctx* ctx, const StreamWork& st, uint32_t* id);
static int register_stream_setup(pov_ctx* ctx, const StreamWork& st, uint32_t* id) {
	auto it = ctx->setup_by_key.find(st.setup_key);
	if(it != ctx->setup_by_key.end()) { *id = it->second; return POV_OK; }
	const int rc = register_stream_setup_slow(ctx, st, id);
	if(rc == POV_OK) ctx->setup_by_key[st.setup_key] = *id;
	return rc;
}

static int register_stream_setup_slow(pov_ctx* ctx, const StreamWork& st, uint32_t* id) {
	SetupAbi a;
	std::string why;
	if(!setup_to_abi(st.setup, a, why)) return pov_fail(ctx, POV_ERR_UNSUPPORTED, "unsupported stream: %s", why.c_str());
	return pov_setup_register(ctx, &a.s, id);
}

extern "C" int pov_ogg_vorbis_decode_memory(pov_ctx* ctx, const uint8_t* data, size_t len, const char* debug_out, pov_decoded* out) {
	if(!ctx || !out || (!data && len)) return POV_ERR_ARG;
	memset(out, 0, sizeof *out);
	std::vector<StreamWork> streams;
	ParseError err;
	if(!parse_ogg_file(data, len, streams, err)) return pov_fail(ctx, POV_ERR_STREAM, "check failed: %s", err.msg.c_str());
	// keep the streams that reached their setup header; all must agree on the channel layout to share one PCM array
	std::vector<const StreamWork*> use;
	for(const StreamWork& st : streams) if(st.have_setup) use.push_back(&st);
	if(use.empty()) return POV_OK;                      // nothing decodable: the reference also returns ok
	for(const StreamWork* st : use)
		if(st->setup.channels != use[0]->setup.channels || st->setup.sample_rate != use[0]->setup.sample_rate)
			return pov_fail(ctx, POV_ERR_UNSUPPORTED, "logical streams with different channel layouts in one file");
	if(debug_out && use.size() != 1) return pov_fail(ctx, POV_ERR_UNSUPPORTED, "debug dump needs exactly one logical stream");
	HostBatch hb;
	for(const StreamWork* st : use) {
		uint32_t id = 0;
		int rc = register_stream_setup(ctx, *st, &id);
		if(rc) return rc;
		hb.append(*st, id);
	}
	const uint32_t C = use[0]->setup.channels;
	out->channels = C; out->sample_rate = use[0]->setup.sample_rate;
	out->audio_packets = (uint32_t) hb.packets.size();
	uint64_t frames = 0;
	for(const StreamWork* st : use) frames += st->frames;
	out->frames = frames;
	if(hb.packets.empty()) return POV_OK;

	pov_batch b = hb.view();
	pov_batch_handle* h = nullptr;
	int rc = pov_batch_upload(ctx, &b, &h);
	if(rc) return rc;
	rc = debug_out ? pov_batch_run_staged(ctx, h) : pov_batch_run(ctx, h);
	std::vector<float> arena(hb.pcm_floats);
	if(!rc) rc = pov_batch_fetch_pcm(ctx, h, arena.data(), arena.size(), 1);
	if(!rc) rc = pov_batch_status(ctx, h, nullptr, 0);
	if(!rc && debug_out) {
		StageHost sg;
		rc = pov_batch_fetch_stage_all(ctx, h, sg);
		std::string werr;
		if(!rc && !write_debug_dump(debug_out, *use[0], *h, 0, sg, arena.data(), werr)) rc = pov_fail(ctx, POV_ERR_ARG, "debug dump: %s", werr.c_str());
	}
	pov_batch_free(ctx, h);
	if(rc) return rc;
	// planar [C][frames] across the (chained) streams
	float* pcm = (float*) malloc(sizeof(float) * std::max<uint64_t>(1, frames * C));
	if(!pcm) return pov_fail(ctx, POV_ERR_ARG, "out of memory");
	uint64_t at = 0;
	for(size_t i = 0; i < use.size(); ++i) {
		const pov_stream& rec = hb.streams[i];
		for(uint32_t c = 0; c < C; ++c)
			memcpy(pcm + (uint64_t) c * frames + at, arena.data() + rec.pcm_base + (uint64_t) c * rec.pcm_frames, sizeof(float) * rec.pcm_frames);
		at += rec.pcm_frames;
	}
	out->pcm = pcm;
	return POV_OK;
}

extern "C" void pov_decoded_free(pov_decoded* d) {
	if(!d) return;
	free(d->pcm);
	d->pcm = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------
// corpus decode: `host_threads` front-end workers parse files into chunks, the calling thread feeds the GPU
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct Chunk {
	uint32_t first_file = 0, n_files = 0;
	std::vector<std::vector<StreamWork>> files;
	std::string error;
};
}  // namespace

cudaError_t pov_checksum_launch(const float* pcm, uint64_t n, double* d_sum, cudaStream_t st, uint64_t* launches);

extern "C" int pov_decode_corpus(pov_ctx* ctx, uint32_t n_files, const uint8_t* const* data, const size_t* len,
                                 uint32_t host_threads, uint64_t* frames_out, uint64_t* total_values_out, double* checksum_out) {
	if(!ctx || (n_files && (!data || !len))) return POV_ERR_ARG;
	cudaSetDevice(ctx->device);
	if(host_threads == 0) host_threads = std::max(1u, std::thread::hardware_concurrency());
	const uint32_t files_per_chunk = 64;
	const uint32_t n_chunks = (n_files + files_per_chunk - 1) / files_per_chunk;
	std::atomic<uint32_t> next_chunk(0);
	std::mutex mu;
	std::condition_variable cv_ready, cv_space;
	std::map<uint32_t, std::unique_ptr<Chunk>> ready;    // by chunk index
	const size_t max_ready = std::max<size_t>(4, 2 * host_threads);
	std::atomic<bool> stop(false);

	auto worker = [&]() {
		for(;;) {
			const uint32_t ci = next_chunk.fetch_add(1);
			if(ci >= n_chunks || stop.load()) return;
			std::unique_ptr<Chunk> ck(new Chunk());
			ck->first_file = ci * files_per_chunk;
			ck->n_files = std::min(files_per_chunk, n_files - ck->first_file);
			ck->files.resize(ck->n_files);
			for(uint32_t i = 0; i < ck->n_files && ck->error.empty(); ++i) {
				ParseError err;
				if(!parse_ogg_file(data[ck->first_file + i], len[ck->first_file + i], ck->files[i], err))
					ck->error = "file " + std::to_string(ck->first_file + i) + ": check failed: " + err.msg;
			}
			std::unique_lock<std::mutex> lk(mu);
			cv_space.wait(lk, [&] { return ready.size() < max_ready || stop.load(); });
			ready[ci] = std::move(ck);
			cv_ready.notify_all();
		}
	};
	std::vector<std::thread> pool;
	for(uint32_t t = 0; t < host_threads; ++t) pool.emplace_back(worker);

	int rc = POV_OK;
	uint64_t total = 0;
	double* d_sum = nullptr;
	double h_sum = 0;
	// Two batch slots: while the GPU transforms chunk i and copies its PCM and status words to pinned host memory, the
	// calling thread assembles and validates chunk i+1. A slot is checked (status words) and reused two chunks later.
	struct Slot {
		pov_ctx* ctx = nullptr;                    // slot 0: the caller's context; slot 1: a sibling context (own stream), so
		                                           // that the pageable uploads of one chunk never wait for the other chunk's work
		double* d_sum = nullptr;
		pov_batch_handle* h = nullptr;
		float* pinned = nullptr; size_t pinned_cap = 0;
		uint32_t* status = nullptr; size_t status_cap = 0;
		cudaEvent_t done = nullptr;
		uint32_t n_packets = 0, first_file = 0;
		bool busy = false;
	} slot[2];
	slot[0].ctx = ctx;
	{
		const char* e = nullptr;
		if(pov_ctx_create(ctx->device, &slot[1].ctx, &e) != POV_OK) rc = pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: %s", e ? e : "sibling context");
	}
	for(auto& sl : slot) {
		if(rc) break;
		if(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess) rc = pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: cudaEventCreate failed");
		else if(cudaMalloc(&sl.d_sum, sizeof(double)) != cudaSuccess || cudaMemsetAsync(sl.d_sum, 0, sizeof(double), sl.ctx->stream) != cudaSuccess)
			rc = pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: cudaMalloc failed");
	}
	(void) d_sum;
	auto retire = [&](Slot& sl) -> int {           // wait for a slot's chunk and turn its status words into the reference's error
		if(!sl.busy) return POV_OK;
		sl.busy = false;
		if(cudaEventSynchronize(sl.done) != cudaSuccess) return pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: chunk failed on the device");
		for(uint32_t p = 0; p < sl.n_packets; ++p)
			if(sl.status[p]) {
				const char* what = (sl.status[p] & POV_PKT_FLOOR_PREDICTED) ? "predicted <= range (hpp:536)"
				                 : (sl.status[p] & POV_PKT_FLOOR_RANGE)     ? "floor[i] < 256 (hpp:587)"
				                                                            : "temp.size() > 0 (hpp:739,748: VQ entry out of range)";
				return pov_fail(ctx, POV_ERR_STREAM, "chunk at file %u, audio packet %u: check failed: %s", sl.first_file, p, what);
			}
		return POV_OK;
	};

	for(uint32_t ci = 0; ci < n_chunks && rc == POV_OK; ++ci) {
		std::unique_ptr<Chunk> ck;
		{
			std::unique_lock<std::mutex> lk(mu);
			cv_ready.wait(lk, [&] { return ready.count(ci) != 0; });
			ck = std::move(ready[ci]);
			ready.erase(ci);
			cv_space.notify_all();
		}
		if(!ck->error.empty()) { rc = pov_fail(ctx, POV_ERR_STREAM, "%s", ck->error.c_str()); break; }
		HostBatch hb;
		for(uint32_t i = 0; i < ck->n_files && rc == POV_OK; ++i) {
			uint64_t frames = 0;
			for(const StreamWork& st : ck->files[i]) {
				if(!st.have_setup) continue;
				uint32_t id = 0;
				rc = register_stream_setup(slot[ci & 1].ctx, st, &id);
				if(rc) { if(slot[ci & 1].ctx != ctx) pov_fail(ctx, rc, "%s", pov_last_error(slot[ci & 1].ctx)); break; }
				hb.append(st, id);
				frames += st.frames;
			}
			if(frames_out) frames_out[ck->first_file + i] = frames;
		}
		if(rc || hb.packets.empty()) continue;
		Slot& sl = slot[ci & 1];
		pov_ctx* cx = sl.ctx;
		rc = retire(sl);                           // the chunk that used this slot two iterations ago
		if(rc) break;
		pov_batch b = hb.view();
		rc = pov_batch_upload(cx, &b, &sl.h);      // pageable sources: staged by the runtime before the call returns
		if(!rc) rc = pov_batch_run(cx, sl.h);
		if(rc && cx != ctx) pov_fail(ctx, rc, "%s", pov_last_error(cx));
		if(!rc) {
			const size_t need = hb.pcm_floats * sizeof(float);
			if(need > sl.pinned_cap) {
				if(sl.pinned) cudaFreeHost(sl.pinned);
				sl.pinned = nullptr;
				sl.pinned_cap = need + need / 4;
				if(cudaMallocHost((void**) &sl.pinned, sl.pinned_cap) != cudaSuccess) { rc = pov_fail(ctx, POV_ERR_CUDA, "cudaMallocHost failed"); break; }
			}
			const size_t sneed = hb.packets.size() * sizeof(uint32_t);
			if(sneed > sl.status_cap) {
				if(sl.status) cudaFreeHost(sl.status);
				sl.status = nullptr;
				sl.status_cap = sneed + sneed / 4;
				if(cudaMallocHost((void**) &sl.status, sl.status_cap) != cudaSuccess) { rc = pov_fail(ctx, POV_ERR_CUDA, "cudaMallocHost failed"); break; }
			}
			cudaError_t e = pov_checksum_launch((const float*) pov_batch_pcm_dev(sl.h), hb.pcm_floats, sl.d_sum, cx->stream, &cx->launches);
			if(e != cudaSuccess) { rc = pov_fail(ctx, POV_ERR_CUDA, "checksum kernel: %s", cudaGetErrorString(e)); break; }
			rc = pov_batch_fetch_pcm(cx, sl.h, sl.pinned, hb.pcm_floats, 0);       // delivery of the PCM to the host (asynchronous)
			if(!rc && cudaMemcpyAsync(sl.status, sl.h->d_status.ptr, sneed, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess)
				rc = pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: status copy failed");
			if(!rc && cudaEventRecord(sl.done, cx->stream) != cudaSuccess) rc = pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: cudaEventRecord failed");
			sl.n_packets = (uint32_t) hb.packets.size(); sl.first_file = ck->first_file; sl.busy = (rc == POV_OK);
		}
		total += hb.pcm_floats;
	}
	for(auto& sl : slot) { const int r2 = retire(sl); if(rc == POV_OK) rc = r2; }
	stop.store(true);
	{ std::lock_guard<std::mutex> lk(mu); cv_space.notify_all(); }
	for(auto& t : pool) t.join();
	for(auto& sl : slot) {
		if(!sl.ctx) continue;
		double part = 0;
		if(rc == POV_OK && sl.d_sum &&
		   (cudaMemcpyAsync(&part, sl.d_sum, sizeof(double), cudaMemcpyDeviceToHost, sl.ctx->stream) != cudaSuccess ||
		    cudaStreamSynchronize(sl.ctx->stream) != cudaSuccess))
			rc = pov_fail(ctx, POV_ERR_CUDA, "pov_decode_corpus: checksum copy failed");
		h_sum += part;
		cudaStreamSynchronize(sl.ctx->stream);
		if(sl.h) pov_batch_free(sl.ctx, sl.h);
		if(sl.pinned) cudaFreeHost(sl.pinned);
		if(sl.status) cudaFreeHost(sl.status);
		if(sl.done) cudaEventDestroy(sl.done);
		if(sl.d_sum) cudaFree(sl.d_sum);
		ctx->launches += (sl.ctx != ctx) ? sl.ctx->launches : 0;
		if(sl.ctx != ctx) pov_ctx_destroy(sl.ctx);
	}
	if(total_values_out) *total_values_out = total;
	if(checksum_out) *checksum_out = h_sum;
	return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// same shape as the reference's C entry point (hpp:1493): decode, discard the PCM, report errors as a string
// ---------------------------------------------------------------------------------------------------------------
extern "C" int pov_ogg_vorbis_full_read_from_memory(const char* data, size_t data_len, const char** error_out) {
	static thread_local char errbuf[512];
	static thread_local pov_ctx* tl_ctx = nullptr;
	if(!tl_ctx) {
		const char* e = nullptr;
		if(pov_ctx_create(0, &tl_ctx, &e) != POV_OK) {
			snprintf(errbuf, sizeof errbuf, "%s", e ? e : "pov_ctx_create failed");
			if(error_out) *error_out = errbuf;
			return 1;
		}
	}
	pov_decoded d;
	const int rc = pov_ogg_vorbis_decode_memory(tl_ctx, (const uint8_t*) data, data_len, nullptr, &d);
	pov_decoded_free(&d);
	if(rc != POV_OK) {
		snprintf(errbuf, sizeof errbuf, "%s", pov_last_error(tl_ctx));
		if(error_out) *error_out = errbuf;
		return 1;
	}
	return 0;
}
