// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Driver that instantiates the UNMODIFIED reference templates where they lie under
// /root/reference (compiled by oracle/Makefile; outputs only into oracle/_ref/).
// It is the gtest/boost-free equivalent of the reference's own CLI main
// (reference unit_tests/test.cc:291-317, "type" 3..6) plus a C entry-point layer so the
// tests and bench.py's cpu_baseline / --impl reference leg can call the reference's
//   Compressor<CanonicalHuffEncoder<> >::compress()            (include/compressor.h:62-73)
//   Decompressor<{,Fast,Table}CanonicalHuffDecoder<> >::decompress()  (include/compressor.h:87-92)
// through ctypes. Nothing from the reference is copied here: this file only #includes it.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <sys/time.h>

#include "compressor.h"
#include "canonical_huff_encoder.h"

using namespace glzip;

static double now_s() {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}

extern "C" {

// file -> file, the reference's compress() (out_path "" means in_path + ".crs2", as the reference does)
int ref_compress_file(const char* in_path, const char* out_path) {
  std::string in(in_path), out(out_path ? out_path : "");
  Compressor<CanonicalHuffEncoder<> > c;
  c.set_file(in, out);
  c.compress();
  c.clear();
  return 0;
}

// kind: 0 = CanonicalHuffDecoder (bit-serial, the semantic oracle), 1 = Fast, 2 = Table<8>
int ref_decompress_file(const char* in_path, const char* out_path, int kind) {
  std::string in(in_path), out(out_path ? out_path : "");
  if (kind == 0) {
    Decompressor<CanonicalHuffDecoder<> > d(in, out);
    d.decompress();
  } else if (kind == 1) {
    Decompressor<FastCanonicalHuffDecoder<> > d(in, out);
    d.decompress();
  } else if (kind == 2) {
    Decompressor<TableCanonicalHuffDecoder<> > d(in, out);
    d.decompress();
  } else {
    return 1;
  }
  return 0;
}

// timed variants: wall seconds of just the reference call (files should live in /dev/shm)
double ref_time_compress_file(const char* in_path, const char* out_path) {
  double t0 = now_s();
  ref_compress_file(in_path, out_path);
  return now_s() - t0;
}

double ref_time_decompress_file(const char* in_path, const char* out_path, int kind) {
  double t0 = now_s();
  ref_decompress_file(in_path, out_path, kind);
  return now_s() - t0;
}

}  // extern "C"

#ifdef REF_DRIVER_MAIN
// glzip_ref <infile> <type> [outfile]   type: 3 compress, 4 decode, 5 fast decode, 6 table decode
int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <infile> <type 3|4|5|6> [outfile]\n", argv[0]);
    return 2;
  }
  int type = atoi(argv[2]);
  const char* out = argc > 3 ? argv[3] : "";
  if (type == 3) return ref_compress_file(argv[1], out);
  if (type >= 4 && type <= 6) return ref_decompress_file(argv[1], out, type - 4);
  fprintf(stderr, "unknown type %d\n", type);
  return 2;
}
#endif
