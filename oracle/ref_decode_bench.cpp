// TEST / BASELINE INFRASTRUCTURE — not on the product path.
//
// Timing driver around the UNMODIFIED reference decoder. It includes the reference's own header from
// /root/reference/src (compile-time include, nothing copied) and drives its public C++ API exactly like the
// reference's demo CLI does (src/main.cpp:53-67): one OggReader + one counting ParseCallbacks per decode,
// fed from memory through OggReader::full_read_from_memory (src/ParseOggVorbis.hpp:1428). The debug sink is
// left at its default (null, src/Callbacks.cpp:95).
//
// usage: ref_decode_bench <file.ogg> <threads> <decodes_per_thread> [repeat_timed]
// prints one JSON line: {"samples": total f32 PCM values produced, "seconds": wall, "threads": T, ...}
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "ParseOggVorbis.hpp"

namespace {

struct CountingCallbacks : ParseCallbacks {
	uint64_t values = 0; // frames * channels
	bool gotPcmData(const std::vector<DataRange<const float>>& channelPcms) override {
		for(const auto& r : channelPcms) values += r.size();
		return true;
	}
};

bool decode_once(const std::vector<uint8_t>& bytes, uint64_t& values, std::string& err) {
	CountingCallbacks cb;
	OggReader reader(cb);
	OkOrError res = reader.full_read_from_memory(bytes.data(), bytes.size());
	if(res.is_error_) { err = res.err_msg_; return false; }
	values += cb.values;
	return true;
}

} // namespace

int main(int argc, char** argv) {
	if(argc < 4) {
		fprintf(stderr, "usage: %s <file.ogg> <threads> <decodes_per_thread>\n", argv[0]);
		return 2;
	}
	const char* fn = argv[1];
	int threads = atoi(argv[2]);
	int per_thread = atoi(argv[3]);
	if(threads < 1) threads = 1;
	if(per_thread < 1) per_thread = 1;

	std::vector<uint8_t> bytes;
	{
		FILE* f = fopen(fn, "rb");
		if(!f) { fprintf(stderr, "cannot open %s\n", fn); return 2; }
		fseek(f, 0, SEEK_END);
		long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		bytes.resize(size_t(sz));
		if(fread(bytes.data(), 1, bytes.size(), f) != bytes.size()) { fclose(f); return 2; }
		fclose(f);
	}

	// warm-up (one decode, single thread) + error check
	{
		uint64_t v = 0; std::string err;
		if(!decode_once(bytes, v, err)) { fprintf(stderr, "reference decode failed: %s\n", err.c_str()); return 1; }
	}

	std::atomic<uint64_t> total(0);
	std::atomic<int> failed(0);
	auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> pool;
	for(int t = 0; t < threads; ++t) {
		pool.emplace_back([&]() {
			uint64_t v = 0; std::string err;
			for(int i = 0; i < per_thread; ++i)
				if(!decode_once(bytes, v, err)) { failed++; break; }
			total += v;
		});
	}
	for(auto& th : pool) th.join();
	auto t1 = std::chrono::steady_clock::now();
	double sec = std::chrono::duration<double>(t1 - t0).count();
	if(failed.load()) { fprintf(stderr, "a worker failed\n"); return 1; }
	printf("{\"samples\": %llu, \"seconds\": %.6f, \"threads\": %d, \"decodes\": %d, \"bytes_per_decode\": %zu}\n",
		(unsigned long long) total.load(), sec, threads, threads * per_thread, bytes.size());
	return 0;
}
