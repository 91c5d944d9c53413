/*
 * jls_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the JPEG-LS (ISO/IEC 14495-1 / ITU-T T.87) scan codec
 * as the reference (team-charls/charls @ 7b9b2da) implements it.  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA path against an independent
 * implementation.  Nothing under charls_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file byte-for-byte against
 *   - the unmodified reference compiled from /root/reference (oracle/_ref/libcharls_ref.so), for encode
 *     (no restart markers, the only thing the reference can encode) and decode (all restart intervals),
 *   - the reference's own fixtures (T.87 Annex-E conformance streams, the five restart fixtures), and
 *   - the committed golden vectors in tests/golden/ (generated from the reference by tools/make_golden.py).
 */
#ifndef JLS_ORACLE_H
#define JLS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes are the negated charls_jpegls_errc values (reference include/charls/public_types.h:28-88) */
#define JLS_ORACLE_ERR_DESTINATION_TOO_SMALL (-3)
#define JLS_ORACLE_ERR_NEED_MORE_DATA (-4)
#define JLS_ORACLE_ERR_INVALID_DATA (-5)
#define JLS_ORACLE_ERR_RESTART_MARKER_NOT_FOUND (-23)
#define JLS_ORACLE_ERR_INVALID_ARGUMENT (-101)

typedef struct jls_scan_params
{
    int32_t width;                /* pixels per line */
    int32_t height;               /* lines */
    int32_t bits_per_sample;      /* 2..16 */
    int32_t component_count;      /* components in THIS scan: 1 (ILV none) or 2..4 (ILV line / sample) */
    int32_t near_lossless;        /* NEAR */
    int32_t interleave_mode;      /* 0 none, 1 line, 2 sample */
    int32_t color_transformation; /* 0 none, 1 HP1, 2 HP2, 3 HP3 (3 components, 8/16 bit, NEAR 0, ILV != none) */
    int32_t threshold1;           /* validated (non-zero) preset coding parameters */
    int32_t threshold2;
    int32_t threshold3;
    int32_t reset_value;
    uint32_t restart_interval; /* in lines (one "MCU row"); 0 = no restart markers */
} jls_scan_params;

/* ISO/IEC 14495-1 C.2.4.1.1.1 defaults; out = {MAXVAL, T1, T2, T3, RESET}. */
void jls_oracle_default_pc_parameters(int32_t maximum_sample_value, int32_t near_lossless, int32_t out[5]);

/*
 * Encodes one scan.  `source` uses the reference's caller layout: `stride` bytes per line, samples are uint8 (bits<=8)
 * or little-endian uint16; for ILV line/sample a line holds `component_count` interleaved samples per pixel.
 * When restart_interval != 0 an RSTm marker (FF D0+m, m cycling 0..7) is written after every interval except the last,
 * and the coder state is reset exactly like the reference DEcoder resets it.
 * Returns the number of bytes written, or a negative JLS_ORACLE_ERR_*.
 */
int64_t jls_oracle_encode_scan(const jls_scan_params* params, const void* source, size_t stride, uint8_t* destination,
                               size_t destination_capacity);

/*
 * Decodes one scan that starts at `source` (first entropy-coded byte after the SOS segment).
 * Returns the number of bytes consumed (the offset of the 0xFF of the marker that ends the scan), or a negative
 * JLS_ORACLE_ERR_*.
 */
int64_t jls_oracle_decode_scan(const jls_scan_params* params, const uint8_t* source, size_t source_size,
                               void* destination, size_t stride);

#ifdef __cplusplus
}
#endif

#endif
