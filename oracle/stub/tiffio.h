/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for libtiff's <tiffio.h> so that the UNMODIFIED reference sources
 * compile in an image without libtiff (oracle/Makefile.ref, target `backend`).  TIFF input is disabled in the
 * oracle build: TIFFOpen() returns NULL (the reference reports "cannot open" and fails the slide); everything else
 * aborts if it is ever reached.  The in-memory (numpy) and NIfTI entry points of the reference do not touch it. */
#ifndef NYXB200_ORACLE_STUB_TIFFIO_H
#define NYXB200_ORACLE_STUB_TIFFIO_H
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct tiff_stub TIFF;
typedef int64_t tmsize_t;
typedef tmsize_t tsize_t;
typedef void* tdata_t;
typedef uint32_t ttile_t;
typedef uint32_t tstrip_t;
typedef uint16_t tsample_t;
typedef uint16_t tdir_t;
typedef uint16_t uint16;
typedef uint32_t uint32;
#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
#define TIFFTAG_BITSPERSAMPLE 258
#define TIFFTAG_COMPRESSION 259
#define TIFFTAG_SAMPLESPERPIXEL 277
#define TIFFTAG_TILEWIDTH 322
#define TIFFTAG_TILELENGTH 323
#define TIFFTAG_SAMPLEFORMAT 339
static inline TIFF* TIFFOpen (const char* name, const char* mode) { (void)name; (void)mode; return NULL; }
static inline void TIFFClose (TIFF* t) { (void)t; }
static inline int TIFFGetField (TIFF* t, uint32_t tag, ...) { (void)t; (void)tag; abort(); return 0; }
static inline int TIFFSetDirectory (TIFF* t, tdir_t d) { (void)t; (void)d; abort(); return 0; }
static inline tdir_t TIFFNumberOfDirectories (TIFF* t) { (void)t; abort(); return 0; }
static inline int TIFFIsTiled (TIFF* t) { (void)t; abort(); return 0; }
static inline tmsize_t TIFFTileSize (TIFF* t) { (void)t; abort(); return 0; }
static inline tmsize_t TIFFScanlineSize (TIFF* t) { (void)t; abort(); return 0; }
static inline tmsize_t TIFFReadTile (TIFF* t, void* buf, uint32_t x, uint32_t y, uint32_t z, uint16_t s) { (void)t; (void)buf; (void)x; (void)y; (void)z; (void)s; abort(); return 0; }
#ifdef __cplusplus
#define NYX_TIFF_DEFAULT(v) = v
#else
#define NYX_TIFF_DEFAULT(v)
#endif
static inline int TIFFReadScanline (TIFF* t, void* buf, uint32_t row, uint16_t s NYX_TIFF_DEFAULT (0)) { (void)t; (void)buf; (void)row; (void)s; abort(); return 0; }
static inline void* _TIFFmalloc (tmsize_t n) { return malloc ((size_t)n); }
static inline void _TIFFfree (void* p) { free (p); }

#ifdef __cplusplus
}
#endif
#endif
