// Mutation fuzzer for the host front end (vorbis_parse.cpp), meant to be built with -fsanitize=address,undefined:
// byte flips / random bytes / truncations of a real file, page CRCs repaired so that the damage reaches the header and
// packet decoders, every input parsed from an exact-size heap block. The parser must reject or accept, never misbehave.
//   fuzz_parser <file.ogg> <iterations> [seed] [hdr]      ("hdr": mutations concentrated in the three header packets)
#include "vorbis_parse.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fstream>
#include <iterator>
#include <random>
// Re-computes page CRCs after mutation so that the corruption reaches the packet decoders.
namespace pov { uint32_t ogg_crc(uint32_t crc, const uint8_t* p, size_t n); }
static void fix_crcs(std::vector<uint8_t>& d) {
	size_t pos = 0;
	while(pos + 27 <= d.size() && memcmp(&d[pos], "OggS", 4) == 0) {
		const uint32_t nseg = d[pos + 26];
		if(pos + 27 + nseg > d.size()) break;
		size_t body = 0; for(uint32_t i = 0; i < nseg; ++i) body += d[pos + 27 + i];
		if(pos + 27 + nseg + body > d.size()) break;
		memset(&d[pos + 22], 0, 4);
		const uint32_t c = pov::ogg_crc(0, &d[pos], 27 + nseg + body);
		memcpy(&d[pos + 22], &c, 4);
		pos += 27 + nseg + body;
	}
}
int main(int argc, char** argv) {
	std::ifstream f(argv[1], std::ios::binary);
	std::vector<uint8_t> orig((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	const int N = atoi(argv[2]); const unsigned seed = argc > 3 ? atoi(argv[3]) : 1;
	std::mt19937 rng(seed);
	int ok = 0, bad = 0;
	for(int it = 0; it < N; ++it) {
		std::vector<uint8_t> d = orig;
		const int nmut = 1 + rng() % 4;
		const bool hdr = argc > 4; const bool body_only = !hdr && rng() % 4 != 0;
		for(int m = 0; m < nmut; ++m) {
			size_t at = hdr ? rng() % 4600 : rng() % d.size();
			if(body_only && at < 5000 && rng() % 2) at = 5000 + rng() % (d.size() - 5000);   // bias towards audio packets
			if(rng() % 3 == 0) d[at] = (uint8_t) rng(); else d[at] ^= (uint8_t) (1u << (rng() % 8));
		}
		if(rng() % 8 == 0) d.resize(rng() % d.size());               // truncation
		if(rng() % 5 != 0) fix_crcs(d);
		// exact-size heap copy so that ASan sees any over-read
		uint8_t* heap = (uint8_t*) malloc(d.size() ? d.size() : 1);
		memcpy(heap, d.data(), d.size());
		std::vector<pov::StreamWork> s; pov::ParseError e;
		if(pov::parse_ogg_file(heap, d.size(), s, e)) ++ok; else ++bad;
		free(heap);
	}
	printf("fuzz: %d parsed, %d rejected\n", ok, bad);
}
