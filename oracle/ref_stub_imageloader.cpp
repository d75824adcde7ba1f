// TEST INFRASTRUCTURE ONLY.  Link-time stub for the five ImageLoader members the reference's
// out-of-core code paths reference (src/nyx/image_loader.h:20-32).  The in-memory hot path never
// calls them; if anything does, abort loudly.
#include <cstdio>
#include <cstdlib>
#include "image_loader.h"

static void nyxref_stub_abort (const char* what)
{
	std::fprintf (stderr, "oracle/_ref: ImageLoader::%s called -- out-of-core IO is not part of the oracle\n", what);
	std::abort();
}

bool ImageLoader::load_tile (size_t, size_t) { nyxref_stub_abort ("load_tile"); return false; }
const std::vector<uint32_t>& ImageLoader::get_int_tile_buffer() { nyxref_stub_abort ("get_int_tile_buffer"); static std::vector<uint32_t> v; return v; }
size_t ImageLoader::get_tile_x (size_t) { nyxref_stub_abort ("get_tile_x"); return 0; }
size_t ImageLoader::get_tile_y (size_t) { nyxref_stub_abort ("get_tile_y"); return 0; }
size_t ImageLoader::get_within_tile_idx (size_t, size_t) { nyxref_stub_abort ("get_within_tile_idx"); return 0; }
