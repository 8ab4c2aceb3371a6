// placeholder, replaced below
#include "api_internal.h"
extern "C" int pov_ogg_vorbis_decode_memory(pov_ctx* ctx, const uint8_t*, size_t, const char*, pov_decoded*) { return pov_fail(ctx, POV_ERR_UNSUPPORTED, "front end not built yet"); }
extern "C" void pov_decoded_free(pov_decoded*) {}
extern "C" int pov_decode_corpus(pov_ctx* ctx, uint32_t, const uint8_t* const*, const size_t*, uint32_t, uint64_t*, uint64_t*, double*) { return pov_fail(ctx, POV_ERR_UNSUPPORTED, "front end not built yet"); }
extern "C" int pov_ogg_vorbis_full_read_from_memory(const char*, size_t, const char**) { return POV_ERR_UNSUPPORTED; }
