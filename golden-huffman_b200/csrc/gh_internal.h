// Internal declarations shared by the host code and the CUDA translation units (not part of the C ABI).
#ifndef GH_INTERNAL_H_
#define GH_INTERNAL_H_

#include <stddef.h>
#include <stdint.h>

#include "gh_codec.h"

namespace gh {

// ---- decode tables -----------------------------------------------------------------------------------------
// The host hands the device only the small canonical tables of the header; the lookup tables are expanded from
// them ON THE DEVICE (dec_build_luts_kernel) so no per-call host work grows with the table size:
//   lut1  2^12 x u16   one codeword per lookup: (symbol << 6) | length, 0 = codeword longer than 12 bits.
//                      The device analogue of TableCanonicalHuffDecoder::lookup_table_
//                      (reference include/canonical_huff_encoder.cc:466-516) with the symbol folded in.
//   lutC  2^13 x u16   as many whole codewords as fit in 13 bits (never the end mark), as ONE addend for the
//                      decoder's cursor word:  (codewords << 10) - total length.  kLutMiss = the first codeword does
//                      not fit or is the end mark. Used where only counts matter.
//   lutW  2^11 x u32x2 up to 4 whole codewords in 11 bits: .x = (8 * codewords << 10) - total length (one addend that
//                      advances the output fill in bits and moves the cursor), .y = the symbols, first in the lowest
//                      byte. .x = kLutMiss when the first codeword does not fit or is the end mark.
//   lutP  2^12 x u32   one or two whole codewords in 12 bits, with the first length kept so a step can stop after
//                      the first: total length | count << 4 | first length << 6 | symbol 1 << 16 | symbol 2 << 24,
//                      0 = first codeword does not fit or is the end mark. Used by the warp-cooperative writer.
constexpr int kLut1Bits = 12;
constexpr int kLutPBits = 12;
#ifndef GH_LUTC_BITS
#define GH_LUTC_BITS 13
#endif
#ifndef GH_LUTW_BITS
#define GH_LUTW_BITS 12
#endif
constexpr int kLutCBits = GH_LUTC_BITS;
constexpr int kLutWBits = GH_LUTW_BITS;
constexpr int kLutWMaxSyms = 4;
constexpr uint32_t kLutMiss = 32;  // see the cursor reader in gh_decode.cu
constexpr int kCurShift = 10;      // table addends count above a 10-bit cursor field

struct DecodeTables {
  uint32_t first_code_lj[34];  // first_code_[len] << (32 - len), "left-justified" as in FastCanonicalHuffDecoder
                               // (reference include/canonical_huff_encoder.cc:437-438); 0xFFFFFFFF = no code of this length
  uint32_t start_pos[34];
  uint16_t symbol[GH_NSYM + 3];  // symbol_[] clamped to 0..256 (unused slots -> 256)
  uint32_t min_len, max_len;
};

// Fills `t` from `code` (header view: symbol_, min/max_len, start_pos_, first_code_).
int build_decode_tables(const gh_code* code, DecodeTables* t);

// ---- encode table: (codeword, length) per byte + the end mark, passed to kernels by value -----------
typedef gh_encode_table EncodeTable;  // include/gh_codec.h: also a member of gh_device_code
int build_encode_table(const gh_code* code, EncodeTable* t);

// gh_decode with a non-zero first-codeword position (the payload of a .crs2 image starts 8-byte aligned, the
// kernels want 16: the image is decoded from the aligned address below with entry_bit = 64).
int decode_full(const uint8_t* d_payload, uint64_t payload_bytes, const gh_code* code, uint32_t entry_bit,
                uint8_t* d_out, uint64_t out_cap, uint64_t* n_out, void* d_workspace, size_t workspace_bytes,
                void* stream);

// gh_encode without the worst-case capacity precondition: stores are bounded by payload_cap instead.
int encode_unchecked(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
                     uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
                     size_t workspace_bytes, void* stream);

// gh_compress_device with the code built on the device: packs d_in behind the header that build_code_kernel wrote to
// the front of d_image (the payload's place follows from d_code->header_bytes, read by the kernel), bounded by
// image_cap bytes. *d_end_bit receives the payload's end bit, counted from the 32-byte boundary below the header's end.
int encode_with_device_code(const uint8_t* d_in, uint64_t n, const gh_device_code* d_code, uint8_t* d_image,
                            uint64_t image_cap, uint64_t* d_end_bit, void* d_workspace, size_t workspace_bytes, void* stream);

}  // namespace gh
#endif
