// Debug dump writer (see debug_dump.cpp).
#ifndef POV_DEBUG_DUMP_H
#define POV_DEBUG_DUMP_H

#include <string>

#include "api_internal.h"
#include "vorbis_parse.h"

namespace pov {

// Writes the dump of ONE logical stream whose packets occupy [first_packet, first_packet + st.packets.size()) of the
// batch behind `h`. `sg` = stage arrays fetched after pov_batch_run_staged, `pcm_planar` = that stream's
// [channels][st.frames] PCM.
bool write_debug_dump(const char* path, const StreamWork& st, const pov_batch_handle& h, uint32_t first_packet,
                      const StageHost& sg, const float* pcm_planar, std::string& err);

}  // namespace pov
#endif
