/* TEST INFRASTRUCTURE ONLY (oracle build). Prototype-only stand-in for <zstd.h>:
 * the image ships libzstd.so.1 (runtime) but not the development header.
 * Declares exactly the streaming symbols the reference's zstdstream.{h,cpp}
 * uses; the implementation is the system's real libzstd. */
#ifndef MALVA_ORACLE_ZSTD_SHIM_H
#define MALVA_ORACLE_ZSTD_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ZSTD_CCtx_s ZSTD_CCtx;
typedef struct ZSTD_DCtx_s ZSTD_DCtx;
typedef ZSTD_CCtx ZSTD_CStream;
typedef ZSTD_DCtx ZSTD_DStream;

typedef struct ZSTD_inBuffer_s {
  const void *src;
  size_t size;
  size_t pos;
} ZSTD_inBuffer;

typedef struct ZSTD_outBuffer_s {
  void *dst;
  size_t size;
  size_t pos;
} ZSTD_outBuffer;

unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
unsigned ZSTD_isFrame(const void *buffer, size_t size);

ZSTD_CStream *ZSTD_createCStream(void);
size_t ZSTD_freeCStream(ZSTD_CStream *zcs);
size_t ZSTD_initCStream(ZSTD_CStream *zcs, int compressionLevel);
size_t ZSTD_compressStream(ZSTD_CStream *zcs, ZSTD_outBuffer *output, ZSTD_inBuffer *input);
size_t ZSTD_flushStream(ZSTD_CStream *zcs, ZSTD_outBuffer *output);
size_t ZSTD_endStream(ZSTD_CStream *zcs, ZSTD_outBuffer *output);
size_t ZSTD_CStreamInSize(void);
size_t ZSTD_CStreamOutSize(void);

ZSTD_DStream *ZSTD_createDStream(void);
size_t ZSTD_freeDStream(ZSTD_DStream *zds);
size_t ZSTD_initDStream(ZSTD_DStream *zds);
size_t ZSTD_decompressStream(ZSTD_DStream *zds, ZSTD_outBuffer *output, ZSTD_inBuffer *input);
size_t ZSTD_DStreamInSize(void);
size_t ZSTD_DStreamOutSize(void);

#ifdef __cplusplus
}
#endif
#endif
