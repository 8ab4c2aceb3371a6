// Writer for the reference's debug dump format "ParseOggVorbis-header-v1" (src/Callbacks.cpp:146-185, 318-324),
// fed from the STAGED device path: every array below was computed by a CUDA kernel and copied back, except the
// coded Y lists and positions, which are the host front-end's own descriptors. The reference's
// tests/compare-debug-out.py (--ourout/--libvorbisout) can diff this file against the reference decoder's dump.
//
// Entry order per packet follows the reference's push sites: hpp:1139-1141, 1164, 518, 560-561, 585, 1171, 1211,
// 1254, 1265, 1270, then "pcm" (hpp:1051) after finish_audio_packet.
#include "debug_dump.h"

#include <stdio.h>
#include <string.h>

namespace pov {

namespace {

enum : uint8_t { DT_F32 = 1, DT_I32 = 2, DT_U32 = 3, DT_U8 = 4, DT_BOOL = 5, DT_I64 = 6, DT_U64 = 7 };   // Callbacks.h:55-63

struct Writer {
	FILE* f;
	void raw(const void* p, uint32_t n) { fwrite(&n, 4, 1, f); if(n) fwrite(p, 1, n, f); }
	void raw(const char* s) { raw(s, (uint32_t) strlen(s)); }
	void typed(const char* key, uint8_t type, uint8_t elem, const void* data, uint32_t count) {
		raw(key);
		raw(&type, 1);
		raw(&elem, 1);
		raw(data, count * elem);
	}
	void entry(const char* name, int channel, uint8_t type, uint8_t elem, const void* data, uint32_t count) {
		typed("entry-name", DT_U8, 1, name, (uint32_t) strlen(name));
		if(channel >= 0) { const uint8_t c = (uint8_t) channel; typed("entry-channel", DT_U8, 1, &c, 1); }
		typed("entry-data", type, elem, data, count);
	}
};

}  // namespace

bool write_debug_dump(const char* path, const StreamWork& st, const pov_batch_handle& h, uint32_t first_packet,
                      const StageHost& sg, const float* pcm_planar, std::string& err) {
	FILE* f = fopen(path, "wb");
	if(!f) { err = std::string("cannot open ") + path; return false; }
	Writer w{f};
	const VorbisSetup& su = st.setup;
	const uint32_t C = su.channels;
	w.raw("ParseOggVorbis-header-v1");
	w.typed("decoder-name", DT_U8, 1, "pov_synth_b200", 14);
	const uint32_t rate = su.sample_rate;
	w.typed("decoder-sample-rate", DT_U32, 4, &rate, 1);
	const uint8_t ch8 = (uint8_t) C;
	w.typed("decoder-num-channels", DT_U8, 1, &ch8, 1);
	for(const Floor1Setup& fl : su.floors) {                               // hpp:1362-1369
		if(fl.type != 1) continue;
		const uint8_t m = (uint8_t) fl.multiplier;
		w.entry("floor1_unpack multiplier", -1, DT_U8, 1, &m, 1);
		std::vector<uint32_t> xs(fl.xs.begin(), fl.xs.end());
		w.entry("floor1_unpack xs", -1, DT_U32, 4, xs.data(), (uint32_t) xs.size());
	}
	w.entry("finish_setup", -1, DT_U8, 1, nullptr, 0);

	std::vector<uint32_t> tmp32;
	const float* residue_base = sg.residue.data();
	for(size_t k = 0; k < st.packets.size(); ++k) {
		const uint32_t p = first_packet + (uint32_t) k;
		const pov_packet& pk = st.packets[k];
		const ModeSetup& mode = su.modes[pk.mode];
		const MappingSetup& mp = su.mappings[mode.mapping];
		const uint32_t n = su.blocksize[mode.blockflag ? 1 : 0], half = n / 2;
		w.entry("start_audio_packet", -1, DT_U8, 1, nullptr, 0);
		const uint64_t abs_pos = st.abs_total_pos[k];
		const int64_t exp_end = st.expected_end[k];
		w.entry("abs_total_pos", -1, DT_U64, 8, &abs_pos, 1);
		w.entry("expected_ending_total_pos", -1, DT_I64, 8, &exp_end, 1);
		uint64_t yo = pk.ys_off;
		const uint64_t so = h.stage_off[p];
		for(uint32_t c = 0; c < C; ++c) {
			const uint8_t fno = mp.submap_floor[mp.mux[c]];
			w.entry("floor_number", (int) c, DT_U8, 1, &fno, 1);
			if(!((pk.floor_used >> c) & 1)) continue;
			const uint32_t posts = (uint32_t) su.floors[fno].xs.size();
			tmp32.assign(st.ys.begin() + yo, st.ys.begin() + yo + posts);
			yo += posts;
			w.entry("floor1 ys", -1, DT_U32, 4, tmp32.data(), posts);
			const size_t slot = ((size_t) p * h.max_channels + c) * POV_MAX_POSTS;
			w.entry("floor1 final_ys", -1, DT_U32, 4, &sg.final_ys[slot], posts);
			w.entry("floor1 step2_flag", -1, DT_BOOL, 1, &sg.step2[slot], posts);
			tmp32.resize(n);
			for(uint32_t i = 0; i < n; ++i) tmp32[i] = sg.floor[so + (size_t) c * n + i];
			w.entry("floor1 floor", -1, DT_U32, 4, tmp32.data(), n);
			w.entry("floor_outputs", (int) c, DT_F32, 4, &sg.floor_out[so + (size_t) c * n], n);
		}
		for(uint32_t c = 0; c < C; ++c) w.entry("after_residue", (int) c, DT_F32, 4, residue_base + h.spec_off[p] + (size_t) c * half, half);
		for(uint32_t c = 0; c < C; ++c) w.entry("after_envelope", (int) c, DT_F32, 4, &sg.env[so / 2 + (size_t) c * half], half);
		for(uint32_t c = 0; c < C; ++c) w.entry("pcm_after_mdct", (int) c, DT_F32, 4, &sg.mdct[so + (size_t) c * n], n);
		w.entry("finish_audio_packet", -1, DT_U8, 1, nullptr, 0);
		if(pk.emit_frames)
			for(uint32_t c = 0; c < C; ++c)
				w.entry("pcm", (int) c, DT_F32, 4, pcm_planar + (size_t) c * st.frames + pk.pcm_off, pk.emit_frames);
	}
	const bool ok = !ferror(f);
	fclose(f);
	if(!ok) err = "write error";
	return ok;
}

}  // namespace pov
