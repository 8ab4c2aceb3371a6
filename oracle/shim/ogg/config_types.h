/* TEST INFRASTRUCTURE. libogg generates this header at configure time; the vendored copy under
 * /root/reference/tests/libvorbis-standalone/ogg/ lacks it (os_types.h:144 includes it on Linux).
 * Five fixed-width typedefs are all it holds. */
#ifndef POV_ORACLE_OGG_CONFIG_TYPES_H
#define POV_ORACLE_OGG_CONFIG_TYPES_H
#include <stdint.h>
typedef int16_t  ogg_int16_t;
typedef uint16_t ogg_uint16_t;
typedef int32_t  ogg_int32_t;
typedef uint32_t ogg_uint32_t;
typedef int64_t  ogg_int64_t;
#endif
