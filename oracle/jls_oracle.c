/*
 * jls_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see jls_oracle.h).
 *
 * A from-scratch, plain-C restatement of the JPEG-LS scan codec as implemented by the reference
 * (team-charls/charls @ 7b9b2da).  Every function cites the reference file:line whose behaviour it
 * restates (paths relative to /root/reference/).  The code is deliberately simple and scalar: it is the
 * checker, never the thing measured or shipped.
 *
 * Parity status: PINNED against the unmodified reference and its fixtures (tests/test_oracle.py).
 */
#include "jls_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------- */
/* Scalar primitives (reference src/jpegls_algorithm.hpp)                                                            */
/* ---------------------------------------------------------------------------------------------------------------- */

/* src/scan_codec.hpp:18-19 -- run-length order table J[0..31] (T.87 A.2.1 step 3). */
static const int J[32] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 9, 10, 11, 12, 13, 14, 15};

/* src/jpegls_algorithm.hpp:14-26 */
static int32_t log2_ceiling(int32_t n)
{
    int32_t x = 0;
    while (n > ((int32_t)1 << x))
        ++x;
    return x;
}

static int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static int32_t iabs(int32_t a) { return a < 0 ? -a : a; }

/* src/jpegls_algorithm.hpp:91-116 */
static int32_t sign_of(int32_t n) { return n < 0 ? -1 : 1; }          /* sign(): (n >> 31) | 1 */
static int32_t bit_wise_sign(int32_t n) { return n < 0 ? -1 : 0; }    /* n >> 31               */
static int32_t apply_sign(int32_t i, int32_t s) { return (s ^ i) - s; }

/* src/jpegls_algorithm.hpp:68-87 (T.87 A.5.2, code segment A.11) */
static int32_t map_error_value(int32_t e) { return e >= 0 ? 2 * e : -2 * e - 1; }
static int32_t unmap_error_value(int32_t m) { return (m & 1) ? -((m + 1) >> 1) : (m >> 1); }

/* src/jpegls_algorithm.hpp:144-161 -- MED predictor (T.87 A.4.2) */
static int32_t predict_med(int32_t ra, int32_t rb, int32_t rc)
{
    const int32_t mn = imin(ra, rb), mx = imax(ra, rb);
    if (rc >= mx)
        return mn;
    if (rc <= mn)
        return mx;
    return ra + rb - rc;
}

/* src/jpegls_preset_coding_parameters.hpp:15-57 */
static int32_t clamp_pc(int32_t i, int32_t j, int32_t maxval) { return (i > maxval || i < j) ? j : i; }

void jls_oracle_default_pc_parameters(int32_t maxval, int32_t near_lossless, int32_t out[5])
{
    int32_t t1, t2, t3;
    if (maxval >= 128)
    {
        const int32_t factor = (imin(maxval, 4095) + 128) / 256;
        t1 = clamp_pc(factor * (3 - 2) + 2 + 3 * near_lossless, near_lossless + 1, maxval);
        t2 = clamp_pc(factor * (7 - 3) + 3 + 5 * near_lossless, t1, maxval);
        t3 = clamp_pc(factor * (21 - 4) + 4 + 7 * near_lossless, t2, maxval);
    }
    else
    {
        const int32_t factor = 256 / (maxval + 1);
        t1 = clamp_pc(imax(2, 3 / factor + 3 * near_lossless), near_lossless + 1, maxval);
        t2 = clamp_pc(imax(3, 7 / factor + 5 * near_lossless), t1, maxval);
        t3 = clamp_pc(imax(4, 21 / factor + 7 * near_lossless), t2, maxval);
    }
    out[0] = maxval;
    out[1] = t1;
    out[2] = t2;
    out[3] = t3;
    out[4] = 64;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Codec state                                                                                                       */
/* ---------------------------------------------------------------------------------------------------------------- */

typedef struct
{
    int32_t a, b, c, n;
} regular_ctx; /* src/regular_mode_context.hpp:140-143 */

typedef struct
{
    int32_t ri_type, a, n, nn;
} run_ctx; /* src/run_mode_context.hpp:118-122 */

typedef struct
{
    /* traits -- src/default_traits.hpp:51-59 (the lossless_traits specialisations are the NEAR=0 case of the same maths) */
    int32_t maxval, near_lossless, range, qbpp, bpp, limit;
    int32_t t1, t2, t3, reset;
    regular_ctx ctx[365];
    run_ctx rctx[2];
    int32_t run_index;
    int error;

    /* bit writer (src/scan_encoder.hpp:75-186) */
    uint8_t* out;
    size_t out_pos, out_cap;
    uint64_t wacc;
    int wbits;
    int w_prev_ff;

    /* bit reader (src/scan_decoder.hpp:250-333) */
    const uint8_t* in;
    size_t in_pos, in_end;
    uint64_t rcache; /* left-aligned */
    int rvalid;
    int r_prev_ff;
} codec;

/* src/scan_codec.hpp:163-174 + src/regular_mode_context.hpp:24-26 + src/jpegls_algorithm.hpp:56-60 */
static void reset_contexts(codec* s)
{
    const int32_t a0 = imax(2, (s->range + 32) / 64);
    for (int i = 0; i < 365; ++i)
    {
        s->ctx[i].a = a0;
        s->ctx[i].b = 0;
        s->ctx[i].c = 0;
        s->ctx[i].n = 1;
    }
    for (int i = 0; i < 2; ++i)
    {
        s->rctx[i].ri_type = i;
        s->rctx[i].a = a0;
        s->rctx[i].n = 1;
        s->rctx[i].nn = 0;
    }
    s->run_index = 0;
}

/* src/make_scan_codec.cpp:98 (codec is always built from 2^bps-1) + src/default_traits.hpp:51-59 */
static void init_codec(codec* s, const jls_scan_params* p)
{
    memset(s, 0, sizeof *s);
    s->maxval = (1 << p->bits_per_sample) - 1;
    s->near_lossless = p->near_lossless;
    s->range = (s->maxval + 2 * s->near_lossless) / (2 * s->near_lossless + 1) + 1;
    s->qbpp = log2_ceiling(s->range);
    s->bpp = log2_ceiling(s->maxval);
    s->limit = 2 * (s->bpp + imax(8, s->bpp));
    s->t1 = p->threshold1;
    s->t2 = p->threshold2;
    s->t3 = p->threshold3;
    s->reset = p->reset_value & 0xFF; /* src/scan_codec.hpp:129: static_cast<uint8_t>(reset_value) */
    reset_contexts(s);
}

/* src/jpegls_algorithm.hpp:173-194 */
static int32_t quantize_gradient(const codec* s, int32_t di)
{
    if (di <= -s->t3)
        return -4;
    if (di <= -s->t2)
        return -3;
    if (di <= -s->t1)
        return -2;
    if (di < -s->near_lossless)
        return -1;
    if (di <= s->near_lossless)
        return 0;
    if (di < s->t1)
        return 1;
    if (di < s->t2)
        return 2;
    if (di < s->t3)
        return 3;
    return 4;
}

static int32_t context_id(const codec* s, int32_t ra, int32_t rb, int32_t rc, int32_t rd)
{
    /* src/jpegls_algorithm.hpp:165-168, call sites src/scan_encoder_impl.hpp:122-124 */
    return (quantize_gradient(s, rd - rb) * 9 + quantize_gradient(s, rb - rc)) * 9 + quantize_gradient(s, rc - ra);
}

/* src/default_traits.hpp:111-117 */
static int32_t correct_prediction(const codec* s, int32_t predicted)
{
    if ((predicted & s->maxval) == predicted)
        return predicted;
    return predicted < 0 ? 0 : s->maxval;
}

/* src/default_traits.hpp:123-139, 157-163 */
static int32_t compute_error_value(const codec* s, int32_t e)
{
    const int32_t d = 2 * s->near_lossless + 1;
    e = e > 0 ? (e + s->near_lossless) / d : -((s->near_lossless - e) / d);
    if (e < 0)
        e += s->range;
    if (e >= (s->range + 1) / 2)
        e -= s->range;
    return e;
}

/* src/default_traits.hpp:77-81, 166-184 */
static int32_t reconstruct(const codec* s, int32_t predicted, int32_t error_value)
{
    const int32_t d = 2 * s->near_lossless + 1;
    /* damaged input can make the product wrap; the reference's int32 arithmetic wraps the same way on the hardware it runs on */
    int32_t v = (int32_t)((uint32_t)predicted + (uint32_t)error_value * (uint32_t)d);
    if (v < -s->near_lossless)
        v += s->range * d;
    else if (v > s->maxval + s->near_lossless)
        v -= s->range * d;
    return correct_prediction(s, v);
}

/* src/regular_mode_context.hpp:99-111 */
static int32_t regular_k(codec* s, const regular_ctx* c)
{
    int32_t k = 0;
    for (; (c->n << k) < c->a && k < 16; ++k)
    {
    }
    if (k == 16)
        s->error = JLS_ORACLE_ERR_INVALID_DATA;
    return k;
}

/* src/regular_mode_context.hpp:36-42 */
static int32_t error_correction(const regular_ctx* c, int32_t k)
{
    if (k != 0)
        return 0;
    return bit_wise_sign(2 * c->b + c->n - 1);
}

/* src/regular_mode_context.hpp:45-94 (T.87 code segments A.12, A.13) */
static void update_regular(codec* s, regular_ctx* c, int32_t e)
{
    c->a += iabs(e);
    c->b += e * (2 * s->near_lossless + 1);
    if (c->a >= 65536 * 256 || iabs(c->b) >= 65536 * 256)
        s->error = JLS_ORACLE_ERR_INVALID_DATA;
    if (c->n == s->reset)
    {
        c->a >>= 1;
        c->b >>= 1; /* arithmetic shift, like the reference */
        c->n >>= 1;
    }
    ++c->n;
    if (c->b + c->n <= 0)
    {
        c->b += c->n;
        if (c->b <= -c->n)
            c->b = -c->n + 1;
        if (c->c > -128)
            --c->c;
    }
    else if (c->b > 0)
    {
        c->b -= c->n;
        if (c->b > 0)
            c->b = 0;
        if (c->c < 127)
            ++c->c;
    }
}

/* src/run_mode_context.hpp:36-61 */
static int32_t run_k(codec* s, const run_ctx* c)
{
    const int32_t temp = c->a + (c->n >> 1) * c->ri_type;
    int32_t n_test = c->n;
    int32_t k = 0;
    for (; n_test < temp; ++k)
    {
        n_test = (int32_t)((uint32_t)n_test << 1); /* may wrap on damaged input, like the reference's */
        if (k > 32)
        {
            s->error = JLS_ORACLE_ERR_INVALID_DATA;
            break;
        }
    }
    return k;
}

/* src/run_mode_context.hpp:102-115 (T.87 code segment A.21) */
static int run_compute_map(const run_ctx* c, int32_t e, int32_t k)
{
    if (k == 0 && e > 0 && 2 * c->nn < c->n)
        return 1;
    if (e < 0 && 2 * c->nn >= c->n)
        return 1;
    if (e < 0 && k != 0)
        return 1;
    return 0;
}

/* src/run_mode_context.hpp:64-82 (T.87 code segment A.23) */
static void update_run(const codec* s, run_ctx* c, int32_t e, int32_t e_mapped)
{
    if (e < 0)
        ++c->nn;
    c->a += (e_mapped + 1 - c->ri_type) >> 1;
    if (c->n == s->reset)
    {
        c->a >>= 1;
        c->n >>= 1;
        c->nn >>= 1;
    }
    ++c->n;
}

/* src/run_mode_context.hpp:84-99 */
static int32_t run_error_value(const run_ctx* c, int32_t temp, int32_t k)
{
    const int map = temp & 1;
    const int32_t e_abs = (temp + map) / 2;
    if ((k != 0 || (2 * c->nn >= c->n)) == map)
        return -e_abs;
    return e_abs;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Bit writer: MSB-first, after a 0xFF byte the next byte carries 7 bits (src/scan_encoder.hpp:75-180, T.87 A.1)     */
/* ---------------------------------------------------------------------------------------------------------------- */

static void bw_emit_ready_bytes(codec* s)
{
    for (;;)
    {
        const int take = s->w_prev_ff ? 7 : 8;
        if (s->wbits < take)
            return;
        const uint8_t b = (uint8_t)((s->wacc >> (s->wbits - take)) & ((1u << take) - 1u));
        s->wbits -= take;
        if (s->out_pos >= s->out_cap)
        {
            s->error = JLS_ORACLE_ERR_DESTINATION_TOO_SMALL;
            return;
        }
        s->out[s->out_pos++] = b;
        s->w_prev_ff = (b == 0xFF);
    }
}

static void bw_put(codec* s, uint32_t value, int count) /* count in [0, 32] */
{
    if (count == 0)
        return;
    const uint64_t mask = count == 32 ? 0xFFFFFFFFull : ((1ull << count) - 1ull);
    s->wacc = (s->wacc << count) | ((uint64_t)value & mask);
    s->wbits += count;
    bw_emit_ready_bytes(s);
}

static void bw_put_zeros(codec* s, int count)
{
    while (count > 0)
    {
        const int n = count > 24 ? 24 : count;
        bw_put(s, 0, n);
        count -= n;
    }
}

/* src/scan_encoder.hpp:103-115 -- pad to a byte with zero bits; a final 0xFF is followed by a (7-bit) zero byte. */
static void bw_end_interval(codec* s)
{
    if (s->wbits > 0)
    {
        const int take = s->w_prev_ff ? 7 : 8;
        bw_put(s, 0, take - s->wbits);
    }
    if (s->w_prev_ff)
        bw_put(s, 0, 7);
}

static void bw_raw_byte(codec* s, uint8_t b)
{
    if (s->out_pos >= s->out_cap)
    {
        s->error = JLS_ORACLE_ERR_DESTINATION_TOO_SMALL;
        return;
    }
    s->out[s->out_pos++] = b;
}

/* src/scan_encoder_core.hpp:69-103 -- limited-length Golomb code (T.87 A.5.3) */
static void encode_mapped_value(codec* s, int32_t k, int32_t mapped, int32_t limit)
{
    const int32_t high = mapped >> k;
    if (high < limit - s->qbpp - 1)
    {
        bw_put_zeros(s, high);
        bw_put(s, 1, 1);
        bw_put(s, (uint32_t)(mapped & ((1 << k) - 1)), k);
        return;
    }
    bw_put_zeros(s, limit - s->qbpp - 1);
    bw_put(s, 1, 1);
    bw_put(s, (uint32_t)((mapped - 1) & ((1 << s->qbpp) - 1)), s->qbpp);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Bit reader (src/scan_decoder.hpp:250-333): stops at a marker (FF followed by a byte with its MSB set)             */
/* ---------------------------------------------------------------------------------------------------------------- */

static void br_fill(codec* s)
{
    while (s->rvalid <= 56)
    {
        if (s->in_pos >= s->in_end)
            return;
        const uint8_t b = s->in[s->in_pos];
        if (b == 0xFF && (s->in_pos + 1 >= s->in_end || (s->in[s->in_pos + 1] & 0x80) != 0))
            return; /* marker or end of buffer */
        if (s->r_prev_ff)
        {
            s->rcache |= (uint64_t)(b & 0x7F) << (64 - 7 - s->rvalid);
            s->rvalid += 7;
        }
        else
        {
            s->rcache |= (uint64_t)b << (64 - 8 - s->rvalid);
            s->rvalid += 8;
        }
        s->r_prev_ff = (b == 0xFF);
        ++s->in_pos;
    }
}

static int32_t br_read(codec* s, int count) /* count in [1, 31] */
{
    if (s->rvalid < count)
    {
        br_fill(s);
        if (s->rvalid < count)
        {
            s->error = JLS_ORACLE_ERR_INVALID_DATA; /* src/scan_decoder.hpp:131-135 */
            return 0;
        }
    }
    const int32_t v = (int32_t)(s->rcache >> (64 - count));
    s->rcache <<= count;
    s->rvalid -= count;
    return v;
}

/* src/scan_decoder.hpp:203-217 */
static int32_t br_read_unary(codec* s)
{
    int32_t zeros = 0;
    for (;;)
    {
        if (s->rvalid == 0)
        {
            br_fill(s);
            if (s->rvalid == 0)
            {
                s->error = JLS_ORACLE_ERR_INVALID_DATA;
                return 0;
            }
        }
        if (s->rcache >> 63)
        {
            s->rcache <<= 1;
            --s->rvalid;
            return zeros;
        }
        s->rcache <<= 1;
        --s->rvalid;
        ++zeros;
    }
}

/* src/scan_decoder.hpp:113-125 */
static int32_t decode_mapped_error_value(codec* s, int32_t k, int32_t limit)
{
    const int32_t unary = br_read_unary(s);
    if (s->error)
        return 0;
    if (unary < limit - s->qbpp - 1)
        return k == 0 ? unary : (unary << k) + br_read(s, k);
    return br_read(s, s->qbpp) + 1;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Regular mode and run mode, encoder side                                                                           */
/* ---------------------------------------------------------------------------------------------------------------- */

/* src/scan_encoder_core.hpp:40-67 */
static int32_t encode_regular(codec* s, int32_t qs, int32_t x, int32_t predicted)
{
    const int32_t sign = bit_wise_sign(qs);
    regular_ctx* c = &s->ctx[apply_sign(qs, sign)];
    const int32_t k = regular_k(s, c);
    const int32_t pv = correct_prediction(s, predicted + apply_sign(c->c, sign));
    const int32_t e = compute_error_value(s, apply_sign(x - pv, sign));
    encode_mapped_value(s, k, map_error_value(error_correction(c, k | s->near_lossless) ^ e), s->limit);
    update_regular(s, c, e);
    return reconstruct(s, pv, apply_sign(e, sign));
}

/* src/scan_encoder.hpp:53-73 */
static void encode_run_pixels(codec* s, int32_t run_length, int end_of_line)
{
    while (run_length >= (1 << J[s->run_index]))
    {
        bw_put(s, 1, 1);
        run_length -= 1 << J[s->run_index];
        if (s->run_index < 31)
            ++s->run_index;
    }
    if (end_of_line)
    {
        if (run_length != 0)
            bw_put(s, 1, 1);
    }
    else
    {
        bw_put(s, (uint32_t)run_length, J[s->run_index] + 1);
    }
}

/* src/scan_encoder_core.hpp:105-116 */
static void encode_run_interruption_error(codec* s, run_ctx* c, int32_t e)
{
    const int32_t k = run_k(s, c);
    const int map = run_compute_map(c, e, k);
    const int32_t e_mapped = 2 * iabs(e) - c->ri_type - map;
    encode_mapped_value(s, k, e_mapped, s->limit - J[s->run_index] - 1);
    update_run(s, c, e, e_mapped);
}

/* src/scan_encoder_core.hpp:118-131 (scalar pixel) */
static int32_t encode_run_interruption_pixel(codec* s, int32_t x, int32_t ra, int32_t rb)
{
    if (iabs(ra - rb) <= s->near_lossless)
    {
        const int32_t e = compute_error_value(s, x - ra);
        encode_run_interruption_error(s, &s->rctx[1], e);
        return reconstruct(s, ra, e);
    }
    const int32_t e = compute_error_value(s, (x - rb) * sign_of(rb - ra));
    encode_run_interruption_error(s, &s->rctx[0], e);
    return reconstruct(s, rb, e * sign_of(rb - ra));
}

/* src/scan_encoder_core.hpp:133-138 (one component of a pair/triplet/quad) */
static int32_t encode_run_interruption_component(codec* s, int32_t x, int32_t ra, int32_t rb)
{
    const int32_t e = compute_error_value(s, sign_of(rb - ra) * (x - rb));
    encode_run_interruption_error(s, &s->rctx[0], e);
    return reconstruct(s, rb, e * sign_of(rb - ra));
}

/* src/scan_encoder_impl.hpp:109-144 + 249-275, scalar line.  cur/prev point at index 0 of a (width + 2) line. */
static void encode_sample_line(codec* s, int32_t* cur, const int32_t* prev, int32_t width)
{
    int32_t index = 1;
    while (index <= width)
    {
        const int32_t ra = cur[index - 1], rc = prev[index - 1], rb = prev[index], rd = prev[index + 1];
        const int32_t qs = context_id(s, ra, rb, rc, rd);
        if (qs != 0)
        {
            cur[index] = encode_regular(s, qs, cur[index], predict_med(ra, rb, rc));
            ++index;
            continue;
        }
        /* run mode */
        const int32_t remain = width - (index - 1);
        int32_t run_length = 0;
        while (iabs(cur[index + run_length] - ra) <= s->near_lossless)
        {
            cur[index + run_length] = ra;
            ++run_length;
            if (run_length == remain)
                break;
        }
        encode_run_pixels(s, run_length, run_length == remain);
        if (run_length == remain)
            return;
        index += run_length;
        cur[index] = encode_run_interruption_pixel(s, cur[index], ra, prev[index]);
        if (s->run_index > 0)
            --s->run_index;
        ++index;
    }
}

/* src/scan_encoder_impl.hpp:147-246 + 249-310, sample-interleaved line with nc components; cur[c]/prev[c] per component. */
static void encode_multi_line(codec* s, int32_t* cur[4], int32_t* prev[4], int32_t width, int nc)
{
    int32_t index = 1;
    while (index <= width)
    {
        int32_t qs[4];
        int all_zero = 1;
        for (int c = 0; c < nc; ++c)
        {
            qs[c] = context_id(s, cur[c][index - 1], prev[c][index], prev[c][index - 1], prev[c][index + 1]);
            if (qs[c] != 0)
                all_zero = 0;
        }
        if (!all_zero)
        {
            for (int c = 0; c < nc; ++c)
                cur[c][index] = encode_regular(s, qs[c], cur[c][index],
                                               predict_med(cur[c][index - 1], prev[c][index], prev[c][index - 1]));
            ++index;
            continue;
        }
        const int32_t remain = width - (index - 1);
        int32_t ra[4];
        for (int c = 0; c < nc; ++c)
            ra[c] = cur[c][index - 1];
        int32_t run_length = 0;
        for (;;)
        {
            int near_all = 1;
            for (int c = 0; c < nc; ++c)
                if (iabs(cur[c][index + run_length] - ra[c]) > s->near_lossless)
                    near_all = 0;
            if (!near_all)
                break;
            for (int c = 0; c < nc; ++c)
                cur[c][index + run_length] = ra[c];
            ++run_length;
            if (run_length == remain)
                break;
        }
        encode_run_pixels(s, run_length, run_length == remain);
        if (run_length == remain)
            return;
        index += run_length;
        for (int c = 0; c < nc; ++c)
            cur[c][index] = encode_run_interruption_component(s, cur[c][index], ra[c], prev[c][index]);
        if (s->run_index > 0)
            --s->run_index;
        ++index;
    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Regular mode and run mode, decoder side                                                                           */
/* ---------------------------------------------------------------------------------------------------------------- */

/* src/scan_decoder_core.hpp:38-69 */
static int32_t decode_regular(codec* s, int32_t qs, int32_t predicted)
{
    const int32_t sign = bit_wise_sign(qs);
    regular_ctx* c = &s->ctx[apply_sign(qs, sign)];
    const int32_t pv = correct_prediction(s, predicted + apply_sign(c->c, sign));
    const int32_t k = regular_k(s, c);
    if (s->error)
        return 0;
    int32_t e = unmap_error_value(decode_mapped_error_value(s, k, s->limit));
    if (iabs(e) > 65535)
        s->error = JLS_ORACLE_ERR_INVALID_DATA;
    if (s->error)
        return 0;
    if (k == 0)
        e ^= error_correction(c, s->near_lossless);
    update_regular(s, c, e);
    return reconstruct(s, pv, apply_sign(e, sign));
}

/* src/scan_decoder_core.hpp:72-80 */
static int32_t decode_run_interruption_error(codec* s, run_ctx* c)
{
    const int32_t k = run_k(s, c);
    if (s->error)
        return 0;
    const int32_t e_mapped = decode_mapped_error_value(s, k, s->limit - J[s->run_index] - 1);
    const int32_t e = run_error_value(c, e_mapped + c->ri_type, k);
    update_run(s, c, e, e_mapped);
    return e;
}

/* src/scan_decoder_impl.hpp:305-337; returns the run length or -1 on error */
static int32_t decode_run_pixels(codec* s, int32_t pixel_count)
{
    int32_t index = 0;
    for (;;)
    {
        const int32_t bit = br_read(s, 1);
        if (s->error)
            return -1;
        if (!bit)
            break;
        const int32_t count = imin(1 << J[s->run_index], pixel_count - index);
        index += count;
        if (count == (1 << J[s->run_index]) && s->run_index < 31)
            ++s->run_index;
        if (index == pixel_count)
            break;
    }
    if (index != pixel_count)
    {
        index += J[s->run_index] > 0 ? br_read(s, J[s->run_index]) : 0;
        if (s->error)
            return -1;
    }
    if (index > pixel_count)
    {
        s->error = JLS_ORACLE_ERR_INVALID_DATA;
        return -1;
    }
    return index;
}

/* src/scan_decoder_impl.hpp:132-159 + 264-281 + src/scan_decoder_core.hpp:83-92 */
static void decode_sample_line(codec* s, int32_t* cur, const int32_t* prev, int32_t width)
{
    int32_t index = 1;
    while (index <= width && !s->error)
    {
        const int32_t ra = cur[index - 1], rc = prev[index - 1], rb = prev[index], rd = prev[index + 1];
        const int32_t qs = context_id(s, ra, rb, rc, rd);
        if (qs != 0)
        {
            cur[index] = decode_regular(s, qs, predict_med(ra, rb, rc));
            ++index;
            continue;
        }
        const int32_t run_length = decode_run_pixels(s, width - (index - 1));
        if (run_length < 0)
            return;
        for (int32_t i = 0; i < run_length; ++i)
            cur[index + i] = ra;
        index += run_length;
        if (index - 1 == width)
            return;
        const int32_t rb2 = prev[index];
        if (iabs(ra - rb2) <= s->near_lossless)
        {
            const int32_t e = decode_run_interruption_error(s, &s->rctx[1]);
            cur[index] = reconstruct(s, ra, e);
        }
        else
        {
            const int32_t e = decode_run_interruption_error(s, &s->rctx[0]);
            cur[index] = reconstruct(s, rb2, e * sign_of(rb2 - ra));
        }
        if (s->run_index > 0)
            --s->run_index;
        ++index;
    }
}

/* src/scan_decoder_impl.hpp:162-261 + 283-303 + src/scan_decoder_core.hpp:94-100 */
static void decode_multi_line(codec* s, int32_t* cur[4], int32_t* prev[4], int32_t width, int nc)
{
    int32_t index = 1;
    while (index <= width && !s->error)
    {
        int32_t qs[4];
        int all_zero = 1;
        for (int c = 0; c < nc; ++c)
        {
            qs[c] = context_id(s, cur[c][index - 1], prev[c][index], prev[c][index - 1], prev[c][index + 1]);
            if (qs[c] != 0)
                all_zero = 0;
        }
        if (!all_zero)
        {
            for (int c = 0; c < nc; ++c)
                cur[c][index] = decode_regular(s, qs[c], predict_med(cur[c][index - 1], prev[c][index], prev[c][index - 1]));
            ++index;
            continue;
        }
        int32_t ra[4];
        for (int c = 0; c < nc; ++c)
            ra[c] = cur[c][index - 1];
        const int32_t run_length = decode_run_pixels(s, width - (index - 1));
        if (run_length < 0)
            return;
        for (int32_t i = 0; i < run_length; ++i)
            for (int c = 0; c < nc; ++c)
                cur[c][index + i] = ra[c];
        index += run_length;
        if (index - 1 == width)
            return;
        for (int c = 0; c < nc; ++c)
        {
            const int32_t rb = prev[c][index];
            const int32_t e = decode_run_interruption_error(s, &s->rctx[0]);
            cur[c][index] = reconstruct(s, rb, e * sign_of(rb - ra[c]));
        }
        if (s->run_index > 0)
            --s->run_index;
        ++index;
    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Caller layout <-> line buffers (src/copy_to_line_buffer.hpp, src/copy_from_line_buffer.hpp, src/color_transform.hpp) */
/* ---------------------------------------------------------------------------------------------------------------- */

static int32_t load_sample(const uint8_t* line, int32_t i, int wide)
{
    return wide ? (int32_t)line[2 * i] | ((int32_t)line[2 * i + 1] << 8) : (int32_t)line[i];
}

static void store_sample(uint8_t* line, int32_t i, int wide, int32_t v)
{
    if (wide)
    {
        line[2 * i] = (uint8_t)(v & 0xFF);
        line[2 * i + 1] = (uint8_t)((v >> 8) & 0xFF);
    }
    else
    {
        line[i] = (uint8_t)v;
    }
}

/* src/copy_to_line_buffer.hpp:26-96 (dispatch), :101-261 (variants); src/color_transform.hpp:27-117 (forward) */
static void load_line(const jls_scan_params* p, const uint8_t* line, int32_t* cur[4])
{
    const int wide = p->bits_per_sample > 8;
    const int nc = p->component_count;
    const int32_t type_mask = wide ? 0xFFFF : 0xFF;       /* static_cast<sample_type>           */
    const int32_t mask = (1 << p->bits_per_sample) - 1;   /* src/scan_encoder.hpp:35             */
    const int32_t range = wide ? 65536 : 256;             /* src/color_transform.hpp:48-50       */
    const int32_t bias = range / 2;
    for (int32_t i = 0; i < p->width; ++i)
    {
        if (nc == 1)
        {
            cur[0][i + 1] = load_sample(line, i, wide) & mask; /* copy_samples / copy_samples_masked */
            continue;
        }
        int32_t v[4];
        for (int c = 0; c < nc; ++c)
            v[c] = load_sample(line, i * nc + c, wide);
        if (nc == 3 && p->color_transformation != 0)
        {
            const int32_t r = v[0], g = v[1], b = v[2];
            switch (p->color_transformation) /* transformed variants do NOT mask (copy_to_line_buffer.hpp:234-246) */
            {
            case 1:
                v[0] = (r - g + bias) & type_mask;
                v[1] = g & type_mask;
                v[2] = (b - g + bias) & type_mask;
                break;
            case 2:
                v[0] = (r - g + bias) & type_mask;
                v[1] = g & type_mask;
                v[2] = (b - ((r + g) / 2) + bias) & type_mask;
                break;
            default: {
                const int32_t v2 = (b - g + bias) & type_mask;
                const int32_t v3 = (r - g + bias) & type_mask;
                v[0] = (g + ((v2 + v3) >> 2) - range / 4) & type_mask;
                v[1] = v2;
                v[2] = v3;
                break;
            }
            }
        }
        else
        {
            for (int c = 0; c < nc; ++c)
                v[c] &= mask;
        }
        for (int c = 0; c < nc; ++c)
            cur[c][i + 1] = v[c];
    }
}

/* src/copy_from_line_buffer.hpp:24-191; src/color_transform.hpp:39-46,70-78,104-112 (inverse) */
static void store_line(const jls_scan_params* p, uint8_t* line, int32_t* cur[4])
{
    const int wide = p->bits_per_sample > 8;
    const int nc = p->component_count;
    const int32_t type_mask = wide ? 0xFFFF : 0xFF;
    const int32_t range = wide ? 65536 : 256;
    const int32_t bias = range / 2;
    for (int32_t i = 0; i < p->width; ++i)
    {
        if (nc == 1)
        {
            store_sample(line, i, wide, cur[0][i + 1]);
            continue;
        }
        int32_t v[4];
        for (int c = 0; c < nc; ++c)
            v[c] = cur[c][i + 1];
        if (nc == 3 && p->color_transformation != 0)
        {
            const int32_t v1 = v[0], v2 = v[1], v3 = v[2];
            switch (p->color_transformation)
            {
            case 1:
                v[0] = (v1 + v2 - bias) & type_mask;
                v[1] = v2 & type_mask;
                v[2] = (v3 + v2 - bias) & type_mask;
                break;
            case 2: {
                const int32_t r = (v1 + v2 - bias) & type_mask;
                v[0] = r;
                v[1] = v2 & type_mask;
                v[2] = (v3 + ((r + (v2 & type_mask)) >> 1) - bias) & type_mask;
                break;
            }
            default: {
                const int32_t g = v1 - ((v3 + v2) >> 2) + range / 4;
                v[0] = (v3 + g - bias) & type_mask;
                v[1] = g & type_mask;
                v[2] = (v2 + g - bias) & type_mask;
                break;
            }
            }
        }
        for (int c = 0; c < nc; ++c)
            store_sample(line, i * nc + c, wide, v[c]);
    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* Scan drivers                                                                                                      */
/* ---------------------------------------------------------------------------------------------------------------- */

typedef struct
{
    int32_t* storage;
    int32_t* line[2][4]; /* [parity][component] -> (width + 2) samples */
    int32_t pixel_stride;
    int nc;
} line_buffers;

static int alloc_lines(line_buffers* lb, int32_t width, int nc)
{
    lb->pixel_stride = width + 2;
    lb->nc = nc;
    lb->storage = (int32_t*)calloc((size_t)2 * (size_t)nc * (size_t)lb->pixel_stride, sizeof(int32_t));
    if (!lb->storage)
        return 0;
    for (int par = 0; par < 2; ++par)
        for (int c = 0; c < nc; ++c)
            lb->line[par][c] = lb->storage + ((size_t)par * (size_t)nc + (size_t)c) * (size_t)lb->pixel_stride;
    return 1;
}

static void zero_lines(line_buffers* lb)
{
    memset(lb->storage, 0, (size_t)2 * (size_t)lb->nc * (size_t)lb->pixel_stride * sizeof(int32_t));
}

static int valid_params(const jls_scan_params* p)
{
    if (!p || p->width < 1 || p->height < 1 || p->bits_per_sample < 2 || p->bits_per_sample > 16)
        return 0;
    if (p->component_count < 1 || p->component_count > 4)
        return 0;
    if (p->interleave_mode == 0 && p->component_count != 1)
        return 0;
    if (p->interleave_mode != 0 && p->component_count < 2)
        return 0;
    if (p->interleave_mode < 0 || p->interleave_mode > 2)
        return 0;
    return 1;
}

/*
 * Line loop: src/scan_encoder_impl.hpp:55-106, restart handling mirrored from the DEcoder
 * (src/scan_decoder_impl.hpp:62-129): after every `restart_interval` lines pad the bit stream, write RSTm,
 * zero both line buffers, re-initialise all contexts and every component's run index.
 */
int64_t jls_oracle_encode_scan(const jls_scan_params* p, const void* source, size_t stride, uint8_t* destination,
                               size_t destination_capacity)
{
    if (!valid_params(p) || !source || !destination)
        return JLS_ORACLE_ERR_INVALID_ARGUMENT;
    codec* s = (codec*)malloc(sizeof(codec));
    line_buffers lb;
    if (!s || !alloc_lines(&lb, p->width, p->component_count))
    {
        free(s);
        return -1;
    }
    init_codec(s, p);
    s->out = destination;
    s->out_cap = destination_capacity;

    const int nc = p->component_count;
    const uint32_t ri = p->restart_interval == 0 ? (uint32_t)p->height : p->restart_interval;
    int32_t run_index[4] = {0, 0, 0, 0};
    uint32_t restart_counter = 0;
    const uint8_t* src = (const uint8_t*)source;

    for (uint32_t line = 0; line < (uint32_t)p->height && !s->error;)
    {
        const uint32_t lines_in_interval = (uint32_t)p->height - line < ri ? (uint32_t)p->height - line : ri;
        for (uint32_t mcu = 0; mcu < lines_in_interval && !s->error; ++mcu, ++line)
        {
            int32_t** cur = lb.line[line & 1];
            int32_t** prev = lb.line[(line & 1) ^ 1];
            load_line(p, src + (size_t)line * stride, cur);
            if (p->interleave_mode == 2)
            {
                s->run_index = run_index[0];
                for (int c = 0; c < nc; ++c)
                {
                    prev[c][p->width + 1] = prev[c][p->width]; /* src/scan_codec.hpp:189-195 */
                    cur[c][0] = prev[c][1];
                }
                encode_multi_line(s, cur, prev, p->width, nc);
                run_index[0] = s->run_index;
            }
            else
            {
                for (int c = 0; c < nc; ++c)
                {
                    s->run_index = run_index[c];
                    prev[c][p->width + 1] = prev[c][p->width];
                    cur[c][0] = prev[c][1];
                    encode_sample_line(s, cur[c], prev[c], p->width);
                    run_index[c] = s->run_index;
                }
            }
        }
        if (line == (uint32_t)p->height)
            break;
        bw_end_interval(s);
        bw_raw_byte(s, 0xFF);
        bw_raw_byte(s, (uint8_t)(0xD0 + restart_counter));
        restart_counter = (restart_counter + 1) % 8;
        s->w_prev_ff = 0;
        memset(run_index, 0, sizeof run_index);
        zero_lines(&lb);
        reset_contexts(s);
    }
    bw_end_interval(s);

    const int64_t result = s->error ? s->error : (int64_t)s->out_pos;
    free(lb.storage);
    free(s);
    return result;
}

/* src/scan_decoder.hpp:237-243, 335-349 */
static void read_restart_marker(codec* s, uint32_t expected_id)
{
    if (s->in_pos >= s->in_end)
    {
        s->error = JLS_ORACLE_ERR_NEED_MORE_DATA;
        return;
    }
    if (s->in[s->in_pos++] != 0xFF)
    {
        s->error = JLS_ORACLE_ERR_RESTART_MARKER_NOT_FOUND;
        return;
    }
    uint8_t v;
    do
    {
        if (s->in_pos >= s->in_end)
        {
            s->error = JLS_ORACLE_ERR_NEED_MORE_DATA;
            return;
        }
        v = s->in[s->in_pos++];
    } while (v == 0xFF);
    if (v != 0xD0 + expected_id)
        s->error = JLS_ORACLE_ERR_RESTART_MARKER_NOT_FOUND;
}

/* After the last symbol of an interval: the stuffed byte that follows a final 0xFF belongs to the interval. */
static void br_finish_interval(codec* s)
{
    if (s->r_prev_ff && s->in_pos < s->in_end && (s->in[s->in_pos] & 0x80) == 0)
    {
        ++s->in_pos;
        s->r_prev_ff = 0;
    }
}

/* src/scan_decoder_impl.hpp:40-129 */
int64_t jls_oracle_decode_scan(const jls_scan_params* p, const uint8_t* source, size_t source_size, void* destination,
                               size_t stride)
{
    if (!valid_params(p) || !source || !destination)
        return JLS_ORACLE_ERR_INVALID_ARGUMENT;
    codec* s = (codec*)malloc(sizeof(codec));
    line_buffers lb;
    if (!s || !alloc_lines(&lb, p->width, p->component_count))
    {
        free(s);
        return -1;
    }
    init_codec(s, p);
    s->in = source;
    s->in_end = source_size;

    const int nc = p->component_count;
    const uint32_t ri = p->restart_interval == 0 ? (uint32_t)p->height : p->restart_interval;
    int32_t run_index[4] = {0, 0, 0, 0};
    uint32_t restart_counter = 0;
    uint8_t* dst = (uint8_t*)destination;

    for (uint32_t line = 0; !s->error;)
    {
        const uint32_t lines_in_interval = (uint32_t)p->height - line < ri ? (uint32_t)p->height - line : ri;
        for (uint32_t mcu = 0; mcu < lines_in_interval && !s->error; ++mcu, ++line)
        {
            int32_t** cur = lb.line[line & 1];
            int32_t** prev = lb.line[(line & 1) ^ 1];
            if (p->interleave_mode == 2)
            {
                s->run_index = run_index[0];
                for (int c = 0; c < nc; ++c)
                {
                    prev[c][p->width + 1] = prev[c][p->width];
                    cur[c][0] = prev[c][1];
                }
                decode_multi_line(s, cur, prev, p->width, nc);
                run_index[0] = s->run_index;
            }
            else
            {
                for (int c = 0; c < nc && !s->error; ++c)
                {
                    s->run_index = run_index[c];
                    prev[c][p->width + 1] = prev[c][p->width];
                    cur[c][0] = prev[c][1];
                    decode_sample_line(s, cur[c], prev[c], p->width);
                    run_index[c] = s->run_index;
                }
            }
            if (!s->error)
                store_line(p, dst + (size_t)line * stride, cur);
        }
        if (s->error || line == (uint32_t)p->height)
            break;

        /* restart: the padding bits are discarded unchecked (src/scan_decoder.hpp:49-56) */
        br_finish_interval(s);
        read_restart_marker(s, restart_counter);
        restart_counter = (restart_counter + 1) % 8;
        s->rcache = 0;
        s->rvalid = 0;
        s->r_prev_ff = 0;
        memset(run_index, 0, sizeof run_index);
        zero_lines(&lb);
        reset_contexts(s);
    }

    if (!s->error)
    {
        /* src/scan_decoder.hpp:71-89: must sit on a marker, left-over (padding) bits must be zero */
        br_finish_interval(s);
        if (s->in_pos < s->in_end && s->in[s->in_pos] != 0xFF)
            br_fill(s); /* the reference's read pointer runs up to one cache ahead of the last symbol */
        if (s->in_pos >= s->in_end)
            s->error = JLS_ORACLE_ERR_NEED_MORE_DATA;
        else if (s->in[s->in_pos] != 0xFF || s->rcache != 0)
            s->error = JLS_ORACLE_ERR_INVALID_DATA;
    }

    const int64_t result = s->error ? s->error : (int64_t)s->in_pos;
    free(lb.storage);
    free(s);
    return result;
}
