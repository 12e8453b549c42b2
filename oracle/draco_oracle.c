/*
 * oracle/draco_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Linear-time plain-C restatement of the attribute-decode hot path of
 * B3zaleel/draco-sharp (a C# port of Google Draco, bitstream v2.2).  Every
 * function cites the reference file:line it follows ("D/" = src/Draco/).
 * The product path (draco_sharp_b200/csrc, libdracob200.so) never links or
 * calls this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, as the checker / CPU baseline.
 *
 * PARITY PINNING.  The reference cannot run here (no .NET in the image) and its
 * own tests hold no golden vector for symbol decode, prediction or
 * dequantisation.  What IS pinned (tests/test_oracle_*.py):
 *   - the reference's own KATs: varint 98 / 1739 and the 9-bit LSB-first value
 *     0b001100010 (tests/Draco.UnitTests/IO/EncoderBufferTests.cs:7-46), the
 *     int -3 <-> uint 4294967293 reinterpret (IO/ConstantsTests.cs:7-21), IntSqrt
 *     0/4/48722615824 (IO/Core/MathUtilitiesTests.cs:7-20);
 *   - the only upstream-produced artefact in the repo,
 *     src/Draco.Examples/Samples/house_04.obj.drc: all nine Raw rANS streams
 *     self-check (0 bytes left, final state == 4*2^precision), the position
 *     attribute's symbols / corrections / quantized ints / floats match the
 *     SHA-256 goldens of SURVEY.md Appendix C, and every dequantised position
 *     lies within half a quantisation step of a `v` line of house_04.obj; its texture
 *     coordinates (TexCoordsPortable predictor, rABS orientation flags) decode to within half a
 *     quantisation step of the `vt` lines.
 * The Tagged symbol scheme is anchored by no upstream artefact (it crashes in the
 * C#, Appendix B-3): for it and for the octahedral transforms parity is
 * "unpinned by reference artefacts" and rests on the mirrored encoder/decoder
 * code plus generator<->oracle round trips.
 *
 * Where the C# throws or corrupts this file follows the
 * Draco bitstream semantics the C# is porting; each such place is marked "B-n".
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off (no -ffast-math: float results
 * must be two separately rounded binary32 operations, as RyuJIT produces).
 */
#include "draco_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* byte reader: BinaryReader semantics (D/IO/DecoderBuffer.cs:51-109)         */
/* ------------------------------------------------------------------------- */
typedef struct {
  const uint8_t *p;
  uint64_t len, pos;
  int err;
} rd_t;

static int rd_need(rd_t *r, uint64_t n) {
  if (r->err) return 0;
  if (r->pos > r->len || r->len - r->pos < n) {
    r->err = ORC_ERR_EOF;
    return 0;
  }
  return 1;
}
static uint8_t rd_u8(rd_t *r) { return rd_need(r, 1) ? r->p[r->pos++] : 0; }
static int8_t rd_i8(rd_t *r) { return (int8_t)rd_u8(r); }
static uint16_t rd_u16(rd_t *r) {
  if (!rd_need(r, 2)) return 0;
  uint16_t v = (uint16_t)(r->p[r->pos] | (r->p[r->pos + 1] << 8));
  r->pos += 2;
  return v;
}
static uint32_t rd_u32(rd_t *r) {
  if (!rd_need(r, 4)) return 0;
  const uint8_t *q = r->p + r->pos;
  r->pos += 4;
  return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
}
static int32_t rd_i32(rd_t *r) { return (int32_t)rd_u32(r); }
static float rd_f32(rd_t *r) {
  uint32_t u = rd_u32(r);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* LEB128 unsigned: D/IO/DecoderBuffer.cs:26-42.  (More than 10 bytes is rejected; the C#
 * would keep shifting with a wrapped byte counter.) */
int orc_varint(const uint8_t *p, uint64_t len, uint64_t *pos, uint64_t *out) {
  uint64_t result = 0;
  unsigned shift = 0;
  for (int i = 0; i < 10; ++i) {
    if (*pos >= len) return ORC_ERR_EOF;
    uint8_t b = p[(*pos)++];
    result |= (uint64_t)(b & 0x7F) << shift;
    if ((b & 0x80) == 0) {
      *out = result;
      return ORC_OK;
    }
    shift += 7;
  }
  return ORC_ERR_EOF;
}
static uint64_t rd_varint(rd_t *r) {
  uint64_t v = 0;
  if (r->err) return 0;
  int e = orc_varint(r->p, r->len, &r->pos, &v);
  if (e) r->err = e;
  return v;
}

/* LSB-first bit reader: D/IO/DecoderBuffer.cs:138-154,177-184 with B-4 applied
 * (full 32-bit assembly; no eager byte).  bitpos counts bits from p[0] bit 0. */
uint32_t orc_read_bits_lsb(const uint8_t *p, uint64_t len, uint64_t *bitpos, int count, int *err) {
  uint32_t value = 0;
  for (int i = 0; i < count; ++i) {
    uint64_t byte = *bitpos >> 3;
    if (byte >= len) {
      if (err) *err = ORC_ERR_EOF;
      return value;
    }
    value |= (uint32_t)((p[byte] >> (*bitpos & 7)) & 1u) << i;
    ++*bitpos;
  }
  return value;
}

/* D/IO/BitUtilities.cs:72-81 */
int32_t orc_zigzag(uint32_t v) {
  int positive = (v & 1u) == 0;
  v >>= 1;
  return positive ? (int32_t)v : (int32_t)(0u - v - 1u);
}
/* D/IO/Constants.cs:183-225 (int -> uint reinterpret used by the wrap transform) */
uint32_t orc_reinterpret_i2u(int32_t v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  return u;
}
/* D/IO/Core/MathUtilities.cs:5-25 */
uint64_t orc_int_sqrt(uint64_t number) {
  if (number == 0) return 0;
  uint64_t act = number, root = 1;
  while (act >= 2) {
    root *= 2;
    act /= 4;
  }
  do {
    root = (root + number / root) / 2;
  } while (root * root > number);
  return root;
}
/* D/IO/Entropy/RAnsSymbolCoding.cs:10-26 */
int orc_rans_precision(int mbl) {
  int p = (3 * mbl) / 2;
  return p < 12 ? 12 : (p > 20 ? 20 : p);
}

/* ------------------------------------------------------------------------- */
/* rANS symbol decoder                                                        */
/* ------------------------------------------------------------------------- */
typedef struct {
  int prec_bits;
  uint32_t prec, l_base;
  uint32_t num_symbols;
  uint32_t *prob, *cum, *lut;
  const uint8_t *buf;
  int64_t off;
  uint32_t state;
} rans_t;

static void rans_free(rans_t *a) {
  free(a->prob);
  free(a->cum);
  free(a->lut);
  memset(a, 0, sizeof *a);
}

/* Table parse: D/IO/Entropy/RAnsSymbolDecoder.cs:12-51; LUT: D/IO/Entropy/RAnsDecoder.cs:69-88 */
static int rans_create(rans_t *a, rd_t *r, int max_bit_length) {
  memset(a, 0, sizeof *a);
  a->prec_bits = orc_rans_precision(max_bit_length);
  a->prec = 1u << a->prec_bits;
  a->l_base = a->prec * 4u;
  uint64_t ns = rd_varint(r);
  if (r->err) return r->err;
  /* every table byte describes at most 64 symbols; anything larger cannot be backed by data */
  if (ns > (r->len - r->pos) * 64u || ns > (1u << 24)) return ORC_ERR_EOF;
  a->num_symbols = (uint32_t)ns;
  if (ns == 0) return ORC_OK;
  a->prob = (uint32_t *)calloc(ns, 4);
  a->cum = (uint32_t *)calloc(ns, 4);
  a->lut = (uint32_t *)calloc(a->prec, 4);
  for (uint32_t i = 0; i < a->num_symbols; ++i) {
    uint8_t pd = rd_u8(r);
    if (r->err) return r->err;
    uint32_t token = pd & 3u;
    if (token == 3) {
      uint32_t offset = (uint32_t)pd >> 2;
      if (i + offset >= a->num_symbols) return ORC_ERR_TABLE; /* :31 */
      for (uint32_t j = 0; j < offset + 1; ++j) a->prob[i + j] = 0;
      i += offset;
    } else {
      uint32_t prob = (uint32_t)pd >> 2;
      for (uint32_t b = 0; b < token; ++b) {
        uint32_t eb = rd_u8(r);
        prob |= eb << (8 * (b + 1) - 2);
      }
      if (r->err) return r->err;
      a->prob[i] = prob;
    }
  }
  uint32_t cum = 0, act = 0;
  for (uint32_t i = 0; i < a->num_symbols; ++i) {
    a->cum[i] = cum;
    /* 64-bit guard: the C# adds in uint and compares, a 2^32 wrap needs prob >= 2^22 which the
     * table encoding cannot express (max 22 bits) but four of them could wrap: reject via 64-bit */
    uint64_t c64 = (uint64_t)cum + a->prob[i];
    if (c64 > a->prec) return ORC_ERR_TABLE; /* RAnsDecoder.cs:80 */
    cum = (uint32_t)c64;
    for (uint32_t j = act; j < cum; ++j) a->lut[j] = i;
    act = cum;
  }
  if (cum != a->prec) return ORC_ERR_TABLE; /* :87 */
  return ORC_OK;
}

/* D/IO/Entropy/RAnsSymbolDecoder.cs:53-59 + D/IO/Entropy/RAnsDecoder.cs:20-54 */
static int rans_start(rans_t *a, rd_t *r, uint64_t *payload_off, uint64_t *payload_len) {
  uint64_t n = rd_varint(r);
  if (r->err) return r->err;
  if (!rd_need(r, n)) return r->err; /* short ReadBytes -> IndexOutOfRange in ReadInit */
  const uint8_t *buf = r->p + r->pos;
  if (payload_off) *payload_off = r->pos;
  if (payload_len) *payload_len = n;
  r->pos += n;
  int64_t offset = (int64_t)n;
  if (offset < 1) return ORC_ERR_RANS_INIT;
  a->buf = buf;
  uint32_t x = (uint32_t)buf[offset - 1] >> 6;
  if (x == 0) {
    a->off = offset - 1;
    a->state = buf[offset - 1] & 0x3Fu;
  } else if (x == 1) {
    if (offset < 2) return ORC_ERR_RANS_INIT;
    a->off = offset - 2;
    a->state = ((uint32_t)buf[offset - 2] | ((uint32_t)buf[offset - 1] << 8)) & 0x3FFFu;
  } else if (x == 2) {
    if (offset < 3) return ORC_ERR_RANS_INIT;
    a->off = offset - 3;
    a->state = ((uint32_t)buf[offset - 3] | ((uint32_t)buf[offset - 2] << 8) | ((uint32_t)buf[offset - 1] << 16)) &
               0x3FFFFFu;
  } else {
    if (offset < 4) return ORC_ERR_RANS_INIT; /* C#: IndexOutOfRangeException */
    a->off = offset - 4;
    a->state = ((uint32_t)buf[offset - 4] | ((uint32_t)buf[offset - 3] << 8) | ((uint32_t)buf[offset - 2] << 16) |
                ((uint32_t)buf[offset - 1] << 24)) &
               0x3FFFFFFFu;
  }
  a->state += a->l_base;
  if (a->state >= a->l_base * 256u) return ORC_ERR_RANS_INIT; /* :53 */
  return ORC_OK;
}

/* D/IO/Entropy/RAnsDecoder.cs:56-67,90-99 */
static inline uint32_t rans_read(rans_t *a) {
  while (a->state < a->l_base && a->off > 0) a->state = a->state * 256u + a->buf[--a->off];
  uint32_t quo = a->state >> a->prec_bits;
  uint32_t rem = a->state & (a->prec - 1u);
  uint32_t s = a->lut[rem];
  a->state = quo * a->prob[s] + rem - a->cum[s];
  return s;
}

/* D/IO/Entropy/SymbolDecoding.cs:7-67 (Tagged with B-3 / B-4 applied) */
int orc_decode_symbols(const uint8_t *p, uint64_t len, uint64_t *pos, uint32_t num_values, uint32_t nc,
                       uint32_t *out, orc_attr *diag) {
  if (num_values == 0) return ORC_OK; /* :9-13 reads nothing */
  rd_t r = {p, len, *pos, 0};
  rans_t a;
  int st = ORC_OK;
  uint8_t scheme = rd_u8(&r);
  if (r.err) return r.err;
  if (diag) diag->scheme = scheme;
  if (scheme == 0) { /* Tagged :30-50 */
    if (diag) {
      diag->max_bit_length = 5;
      diag->table_off = r.pos;
    }
    st = rans_create(&a, &r, 5);
    if (diag) {
      diag->precision = a.prec_bits;
      diag->table_symbols = a.num_symbols;
    }
    if (!st && a.num_symbols == 0) st = ORC_ERR_NUM_SYMBOLS; /* :36 (ReadInit would also fail on null tables) */
    if (!st) st = rans_start(&a, &r, diag ? &diag->payload_off : NULL, diag ? &diag->payload_len : NULL);
    if (st) {
      rans_free(&a);
      return st;
    }
    const uint8_t *bits = p + r.pos;
    uint64_t bits_len = len - r.pos, bitpos = 0;
    uint32_t vid = 0;
    for (uint32_t i = 0; i < num_values; i += nc) {
      uint32_t bit_length = rans_read(&a) & 0xFFu; /* (byte) cast :41 */
      if (bit_length > 32) {
        st = ORC_ERR_TAG;
        break;
      }
      for (uint32_t j = 0; j < nc && vid < num_values; ++j) {
        int e = 0;
        out[vid++] = orc_read_bits_lsb(bits, bits_len, &bitpos, (int)bit_length, &e);
        if (e) st = e;
      }
      if (st) break;
    }
    if (diag) {
      diag->final_state = a.state;
      diag->leftover = (uint64_t)a.off;
      diag->bits_off = r.pos;
      diag->bits_len = (bitpos + 7) / 8;
    }
    r.pos += (bitpos + 7) / 8; /* EndBitDecoding: ceil(bits/8) bytes (B-4) */
    rans_free(&a);
    if (st) return st;
  } else if (scheme == 1) { /* Raw :52-67 */
    uint8_t mbl = rd_u8(&r);
    if (r.err) return r.err;
    if (mbl < 1 || mbl > 18) return ORC_ERR_BITLEN;
    if (diag) {
      diag->max_bit_length = mbl;
      diag->table_off = r.pos;
    }
    st = rans_create(&a, &r, mbl);
    if (diag) {
      diag->precision = a.prec_bits;
      diag->table_symbols = a.num_symbols;
    }
    if (!st && a.num_symbols == 0) st = ORC_ERR_NUM_SYMBOLS; /* :59 */
    if (!st) st = rans_start(&a, &r, diag ? &diag->payload_off : NULL, diag ? &diag->payload_len : NULL);
    if (st) {
      rans_free(&a);
      return st;
    }
    for (uint32_t i = 0; i < num_values; ++i) out[i] = rans_read(&a);
    if (diag) {
      diag->final_state = a.state;
      diag->leftover = (uint64_t)a.off;
    }
    rans_free(&a);
  } else {
    return ORC_ERR_SCHEME; /* :26 */
  }
  *pos = r.pos;
  return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* prediction schemes + transforms                                            */
/* ------------------------------------------------------------------------- */

/* D/IO/Attributes/PredictionSchemes/PredictionSchemeWrapTransform.cs:67-86 (clamp) and
 * PredictionSchemeWrapDecodingTransform.cs:46-67 (mod-2^32 add, one +-max_diff) */
static inline int32_t wrap_original(int32_t pred, int32_t corr, int32_t mn, int32_t mx, int32_t max_diff) {
  if (pred > mx)
    pred = mx;
  else if (pred < mn)
    pred = mn;
  int32_t o = (int32_t)((uint32_t)pred + (uint32_t)corr);
  if (o > mx)
    o = (int32_t)((uint32_t)o - (uint32_t)max_diff);
  else if (o < mn)
    o = (int32_t)((uint32_t)o + (uint32_t)max_diff);
  return o;
}

/* D/IO/Attributes/PredictionSchemes/PredictionSchemeDeltaDecoder.cs:23-37 with the wrap transform */
void orc_delta_wrap(const int32_t *corr, uint32_t n, int nc, int32_t mn, int32_t mx, int32_t *out) {
  int32_t max_diff = (int32_t)(1u + (uint32_t)mx - (uint32_t)mn); /* WrapTransform.cs:88-92 */
  for (int c = 0; c < nc && n > 0; ++c) out[c] = wrap_original(0, corr[c], mn, mx, max_diff); /* :30 */
  for (uint64_t i = (uint64_t)nc; i < (uint64_t)n * nc; ++i)
    out[i] = wrap_original(out[i - nc], corr[i], mn, mx, max_diff); /* :32-35 */
}

/* D/IO/Attributes/PredictionSchemes/MeshPredictionSchemeParallelogramDecoder.cs:29-89 */
int orc_parallelogram_wrap(const int32_t *corr, uint32_t n, int nc, int32_t mn, int32_t mx,
                           const orc_mesh_maps *m, int32_t *out) {
  int32_t max_diff = (int32_t)(1u + (uint32_t)mx - (uint32_t)mn);
  if (n == 0) return ORC_OK;
  if (!m || m->n_entries < n) return ORC_ERR_MAPS;
  for (int c = 0; c < nc; ++c) out[c] = wrap_original(0, corr[c], mn, mx, max_diff); /* :36 */
  for (uint32_t p = 1; p < n; ++p) {                                                 /* :38 */
    uint32_t corner = m->data_to_corner[p];
    uint64_t dst = (uint64_t)p * nc;
    int used = 0;
    if (corner != 0xFFFFFFFFu && corner < m->n_corners) {
      uint32_t oc = m->opposite[corner]; /* :66 */
      if (oc != 0xFFFFFFFFu) {
        if (oc >= m->n_corners) return ORC_ERR_MAPS;
        uint32_t nx = (oc % 3u == 2u) ? oc - 2u : oc + 1u; /* CornerTable.Next :64-67 */
        uint32_t pv = (oc % 3u == 0u) ? oc + 2u : oc - 1u; /* CornerTable.Previous :69-72 */
        uint32_t v_o = m->corner_to_vertex[oc], v_n = m->corner_to_vertex[nx], v_p = m->corner_to_vertex[pv];
        if (v_o >= m->n_vertices || v_n >= m->n_vertices || v_p >= m->n_vertices) return ORC_ERR_MAPS;
        int32_t e_o = m->vertex_to_data[v_o], e_n = m->vertex_to_data[v_n], e_p = m->vertex_to_data[v_p]; /* :56-59 */
        if (e_o < (int32_t)p && e_n < (int32_t)p && e_p < (int32_t)p) {                                   /* :75 */
          if (e_o < 0 || e_n < 0 || e_p < 0) return ORC_ERR_MAPS; /* C#: IndexOutOfRangeException */
          for (int c = 0; c < nc; ++c) {
            int32_t pred = (int32_t)((uint32_t)out[(uint64_t)e_n * nc + c] + (uint32_t)out[(uint64_t)e_p * nc + c] -
                                     (uint32_t)out[(uint64_t)e_o * nc + c]); /* :84 */
            out[dst + c] = wrap_original(pred, corr[dst + c], mn, mx, max_diff);
          }
          used = 1;
        }
      }
    }
    if (!used) {
      uint64_t src = (uint64_t)(p - 1) * nc; /* :49-50 */
      for (int c = 0; c < nc; ++c) out[dst + c] = wrap_original(out[src + c], corr[dst + c], mn, mx, max_diff);
    }
  }
  return ORC_OK;
}

/* One parallelogram of entry p at `corner`: MeshPredictionSchemeParallelogramDecoder.TryComputeParallelogramPrediction
 * (:62-89) without the arithmetic -- the three operand entries, or 0 when the corner has no usable parallelogram. */
static int para_entries(const orc_mesh_maps *m, uint32_t p, uint32_t corner, int32_t e[3], int *bad) {
  if (corner >= m->n_corners) { *bad = 1; return 0; }
  uint32_t oc = m->opposite[corner]; /* :66 */
  if (oc == 0xFFFFFFFFu) return 0;
  if (oc >= m->n_corners) { *bad = 1; return 0; }
  uint32_t nx = (oc % 3u == 2u) ? oc - 2u : oc + 1u;
  uint32_t pv = (oc % 3u == 0u) ? oc + 2u : oc - 1u;
  uint32_t v_o = m->corner_to_vertex[oc], v_n = m->corner_to_vertex[nx], v_p = m->corner_to_vertex[pv];
  if (v_o >= m->n_vertices || v_n >= m->n_vertices || v_p >= m->n_vertices) { *bad = 1; return 0; }
  e[0] = m->vertex_to_data[v_o]; e[1] = m->vertex_to_data[v_n]; e[2] = m->vertex_to_data[v_p]; /* :56-59 */
  if (e[0] < (int32_t)p && e[1] < (int32_t)p && e[2] < (int32_t)p) {                            /* :75 */
    if (e[0] < 0 || e[1] < 0 || e[2] < 0) { *bad = 1; return 0; }
    return 1;
  }
  return 0;
}
static uint32_t m_next(uint32_t c) { return c == 0xFFFFFFFFu ? c : ((c % 3u == 2u) ? c - 2u : c + 1u); }
static uint32_t m_prev(uint32_t c) { return c == 0xFFFFFFFFu ? c : ((c % 3u == 0u) ? c + 2u : c - 1u); }
static uint32_t m_opp(const orc_mesh_maps *m, uint32_t c, int *bad) {
  if (c == 0xFFFFFFFFu) return c;
  if (c >= m->n_corners) { *bad = 1; return 0xFFFFFFFFu; }
  return m->opposite[c];
}

/* MeshPredictionSchemeConstrainedMultiParallelogramDecoder.ComputeOriginalValues (:32-116) with the wrap transform, in
 * the bitstream's semantics where the C# is defective (`predictedValues[j][j]` is never filled,
 * the crease flags are never decoded): every parallelogram found while swinging left, then right, around the entry's
 * vertex (at most four, :60-80) keeps its own prediction; the flags of context (count - 1) say which of them are
 * crease edges (:89-101); the prediction is the truncated integer mean of the others (:112), or entry p-1 when none
 * is left (:105-108).  crease[c] / n_crease[c]: the decoded flag sequence of context c (DecodeTransformData :125-139). */
int orc_cmp_wrap(const int32_t *corr, uint32_t n, int nc, int32_t mn, int32_t mx, const orc_mesh_maps *m,
                 uint8_t *const crease[4], const uint32_t n_crease[4], int32_t *out) {
  int32_t max_diff = (int32_t)(1u + (uint32_t)mx - (uint32_t)mn);
  if (n == 0) return ORC_OK;
  if (!m || m->n_entries < n) return ORC_ERR_MAPS;
  if (nc > 16) return ORC_ERR_UNSUPPORTED;
  uint32_t pos[4] = {0, 0, 0, 0};
  for (int c = 0; c < nc; ++c) out[c] = wrap_original(0, corr[c], mn, mx, max_diff); /* :43 */
  for (uint32_t p = 1; p < n; ++p) {                                                 /* :45 */
    const uint32_t start = m->data_to_corner[p];
    uint32_t corner = start;
    int32_t pe[4][3];
    int np = 0, first_pass = 1, bad = 0;
    uint64_t guard = 0;
    while (corner != 0xFFFFFFFFu) { /* :52-80 */
      if (++guard > (uint64_t)m->n_corners + 2) return ORC_ERR_MAPS; /* not a corner table: the swing never closes */
      if (para_entries(m, p, corner, pe[np], &bad)) {
        if (++np == 4) break; /* Constants.ConstrainedMultiParallelogramMaxNumParallelograms */
      }
      if (bad) return ORC_ERR_MAPS;
      corner = first_pass ? m_next(m_opp(m, m_next(corner), &bad)) : m_prev(m_opp(m, m_prev(corner), &bad)); /* SwingLeft / SwingRight */
      if (bad) return ORC_ERR_MAPS;
      if (corner == start) break;
      if (corner == 0xFFFFFFFFu && first_pass) {
        first_pass = 0;
        corner = m_prev(m_opp(m, m_prev(start), &bad));
        if (bad) return ORC_ERR_MAPS;
      }
    }
    int32_t sum[16];
    int used = 0;
    for (int c = 0; c < nc; ++c) sum[c] = 0;
    for (int i = 0; i < np; ++i) { /* :89-101 */
      const int ctx = np - 1;
      const uint32_t at = pos[ctx]++;
      if (at >= n_crease[ctx]) return ORC_ERR_PRED; /* :93 */
      if (!crease[ctx][at]) {
        ++used;
        for (int c = 0; c < nc; ++c) {
          int32_t pred = (int32_t)((uint32_t)out[(uint64_t)pe[i][1] * nc + c] + (uint32_t)out[(uint64_t)pe[i][2] * nc + c] -
                                   (uint32_t)out[(uint64_t)pe[i][0] * nc + c]); /* ParallelogramDecoder :84 */
          sum[c] = (int32_t)((uint32_t)sum[c] + (uint32_t)pred);          /* AddAsUnsigned :98 */
        }
      }
    }
    uint64_t dst = (uint64_t)p * nc;
    if (used == 0) {
      uint64_t src = (uint64_t)(p - 1) * nc; /* :105-108 */
      for (int c = 0; c < nc; ++c) out[dst + c] = wrap_original(out[src + c], corr[dst + c], mn, mx, max_diff);
    } else {
      for (int c = 0; c < nc; ++c) out[dst + c] = wrap_original(sum[c] / used, corr[dst + c], mn, mx, max_diff); /* :112-114 */
    }
  }
  return ORC_OK;
}

/* MeshPredictionSchemeTexCoordsPortableDecoder.ComputeOriginalValues (:49-66) over
 * MeshPredictionSchemeTexCoordsPortablePredictor.ComputePredictedValue (:52-150) and the wrap transform.
 * uv: maps of the attribute's own decoder; pos_q / pm: quantized positions (portable parent attribute) and the maps of
 * ITS decoder -- GetPositionForEntryId (:34-39) goes entry -> point -> position value, i.e. through the corner the
 * entry was first reached at.  orient[0..n_or): flags in decoding order; the predictor pops them from the BACK (:127).
 * Where the C# would throw (assertions, negative indices) the stream is invalid: ORC_ERR_PRED / ORC_ERR_MAPS. */
int orc_texcoords_portable_wrap(const int32_t *corr, uint32_t n, int32_t mn, int32_t mx, const orc_mesh_maps *uv,
                                const int32_t *pos_q, uint32_t n_pos, const orc_mesh_maps *pm, const uint8_t *orient,
                                uint32_t n_or, int32_t *out) {
  int32_t max_diff = (int32_t)(1u + (uint32_t)mx - (uint32_t)mn);
  if (n == 0) return ORC_OK;
  if (!uv || !pm || uv->n_entries < n) return ORC_ERR_MAPS;
  uint32_t left = n_or;
  for (uint32_t p = 0; p < n; ++p) {
    uint32_t corner = uv->data_to_corner[p];
    if (corner == 0xFFFFFFFFu || corner >= uv->n_corners) return ORC_ERR_MAPS;
    uint32_t nx = (corner % 3u == 2u) ? corner - 2u : corner + 1u;
    uint32_t pv = (corner % 3u == 0u) ? corner + 2u : corner - 1u;
    uint32_t v_n = uv->corner_to_vertex[nx], v_p = uv->corner_to_vertex[pv];
    if (v_n >= uv->n_vertices || v_p >= uv->n_vertices) return ORC_ERR_MAPS;
    int32_t nd = uv->vertex_to_data[v_n], pd = uv->vertex_to_data[v_p]; /* :58-59 */
    int64_t pred[2] = {0, 0};
    int done = 0;
    if (pd < (int32_t)p && nd < (int32_t)p) { /* :61 */
      if (pd < 0 || nd < 0) return ORC_ERR_MAPS;
      int64_t n_uv[2] = {out[2ull * nd], out[2ull * nd + 1]}, p_uv[2] = {out[2ull * pd], out[2ull * pd + 1]};
      if (p_uv[0] == n_uv[0] && p_uv[1] == n_uv[1]) { /* :66-71 */
        pred[0] = p_uv[0];
        pred[1] = p_uv[1];
        done = 1;
      } else {
        int64_t P[3][3]; /* tip, next, prev */
        const uint32_t ids[3] = {p, (uint32_t)nd, (uint32_t)pd};
        for (int k = 0; k < 3; ++k) {
          uint32_t c = uv->data_to_corner[ids[k]];
          if (c >= pm->n_corners) return ORC_ERR_MAPS;
          uint32_t v = pm->corner_to_vertex[c];
          if (v >= pm->n_vertices) return ORC_ERR_MAPS;
          int32_t e = pm->vertex_to_data[v];
          if (e < 0 || (uint32_t)e >= n_pos) return ORC_ERR_MAPS;
          for (int j = 0; j < 3; ++j) P[k][j] = pos_q[3ull * (uint32_t)e + j];
        }
        int64_t pn[3], cn[3], pn2 = 0, cdp = 0;
        for (int j = 0; j < 3; ++j) {
          pn[j] = P[2][j] - P[1][j];
          cn[j] = P[0][j] - P[1][j];
          pn2 += pn[j] * pn[j];
          cdp += pn[j] * cn[j];
        }
        if (pn2 != 0) { /* :78 */
          int64_t pn_uv[2] = {p_uv[0] - n_uv[0], p_uv[1] - n_uv[1]};
          int64_t a0 = n_uv[0] < 0 ? -n_uv[0] : n_uv[0], a1 = n_uv[1] < 0 ? -n_uv[1] : n_uv[1];
          if ((a0 > a1 ? a0 : a1) > INT64_MAX / pn2) return ORC_ERR_PRED; /* :85 */
          int64_t b0 = pn_uv[0] < 0 ? -pn_uv[0] : pn_uv[0], b1 = pn_uv[1] < 0 ? -pn_uv[1] : pn_uv[1];
          if (cdp > INT64_MAX / (b0 > b1 ? b0 : b1)) return ORC_ERR_PRED; /* :87 */
          int64_t x_uv[2] = {n_uv[0] * pn2 + cdp * pn_uv[0], n_uv[1] * pn2 + cdp * pn_uv[1]};
          int64_t m = 0;
          for (int j = 0; j < 3; ++j) {
            int64_t t = pn[j] < 0 ? -pn[j] : pn[j];
            if (t > m) m = t;
          }
          if (cdp > INT64_MAX / m) return ORC_ERR_PRED; /* :90 */
          int64_t cx2 = 0;
          for (int j = 0; j < 3; ++j) {
            int64_t xp = P[1][j] + (cdp * pn[j]) / pn2; /* :91, truncating */
            int64_t dlt = P[0][j] - xp;
            cx2 += dlt * dlt;
          }
          int64_t nrm = (int64_t)orc_int_sqrt((uint64_t)(cx2 * pn2)); /* :94 */
          int64_t cx_uv[2] = {pn_uv[1] * nrm, -pn_uv[0] * nrm};
          if (left == 0) return ORC_ERR_PRED; /* :125 */
          int o = orient[--left];             /* :126-127: Last() + PopBack() */
          for (int j = 0; j < 2; ++j) pred[j] = (o ? x_uv[j] + cx_uv[j] : x_uv[j] - cx_uv[j]) / pn2; /* :128 */
          done = 1;
        }
      }
    }
    if (!done) { /* :135-160 */
      int64_t off;
      int have = 1;
      off = 0;
      if (pd < (int32_t)p) off = (int64_t)pd * 2;
      if (nd < (int32_t)p) {
        off = (int64_t)nd * 2;
      } else if (p > 0) {
        off = (int64_t)(p - 1) * 2;
      } else {
        have = 0;
      }
      if (have) {
        if (off < 0) return ORC_ERR_MAPS;
        pred[0] = out[off];
        pred[1] = out[off + 1];
      }
    }
    for (int c = 0; c < 2; ++c)
      out[2ull * p + c] = wrap_original((int32_t)pred[c], corr[2ull * p + c], mn, mx, max_diff);
  }
  return ORC_OK;
}

/* Octahedron tool box: D/IO/Attributes/OctahedronToolBox.cs:13-21,144-212 */
typedef struct {
  int32_t bits, max_q, max_value, center;
} octbox_t;
static void octbox_set(octbox_t *t, int bits) {
  t->bits = bits;
  t->max_q = (int32_t)((1u << bits) - 1u);
  t->max_value = t->max_q - 1;
  t->center = t->max_value / 2;
}
static inline int32_t iabs32(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }
static inline int oct_in_diamond(const octbox_t *t, int32_t s, int32_t tt) { /* :144-150 (asserts dropped, B-9) */
  uint32_t st = (uint32_t)iabs32(s) + (uint32_t)iabs32(tt);
  return st <= (uint32_t)t->center;
}
static inline void oct_invert_diamond(const octbox_t *t, int32_t *s, int32_t *tt) { /* :152-194 */
  int32_t sign_s, sign_t;
  if (*s >= 0 && *tt >= 0) {
    sign_s = 1;
    sign_t = 1;
  } else if (*s <= 0 && *tt <= 0) {
    sign_s = -1;
    sign_t = -1;
  } else {
    sign_s = (*s > 0) ? 1 : -1;
    sign_t = (*tt > 0) ? 1 : -1;
  }
  int32_t corner_s = sign_s * t->center, corner_t = sign_t * t->center;
  /* all arithmetic mod 2^32 (C# int, unchecked) */
  int32_t us = (int32_t)((uint32_t)*s + (uint32_t)*s - (uint32_t)corner_s);
  int32_t ut = (int32_t)((uint32_t)*tt + (uint32_t)*tt - (uint32_t)corner_t);
  int32_t temp = us;
  if (sign_s * sign_t >= 0) {
    us = (int32_t)(0u - (uint32_t)ut);
    ut = (int32_t)(0u - (uint32_t)temp);
  } else {
    us = ut;
    ut = temp;
  }
  us = (int32_t)((uint32_t)us + (uint32_t)corner_s);
  ut = (int32_t)((uint32_t)ut + (uint32_t)corner_t);
  *s = us / 2; /* truncating division, as C# */
  *tt = ut / 2;
}
static inline int32_t oct_mod_max(const octbox_t *t, int32_t x) { /* :205-212 */
  if (x > t->center) return (int32_t)((uint32_t)x - (uint32_t)t->max_q);
  return x < -t->center ? (int32_t)((uint32_t)x + (uint32_t)t->max_q) : x;
}
/* ...CanonicalizedTransform.cs:43-89 */
static inline int oct_rotation_count(int32_t sx, int32_t sy) {
  if (sx == 0) return sy == 0 ? 0 : (sy > 0 ? 3 : 1);
  if (sx > 0) return sy >= 0 ? 2 : 1;
  return sy <= 0 ? 0 : 3;
}
static inline void oct_rotate(int32_t *p0, int32_t *p1, int rot) {
  int32_t a = *p0, b = *p1;
  switch (rot) {
    case 1: *p0 = b; *p1 = (int32_t)(0u - (uint32_t)a); break;
    case 2: *p0 = (int32_t)(0u - (uint32_t)a); *p1 = (int32_t)(0u - (uint32_t)b); break;
    case 3: *p0 = (int32_t)(0u - (uint32_t)b); *p1 = a; break;
    default: break;
  }
}
/* PredictionSchemeNormalOctahedronCanonicalizedDecodingTransform.cs:48-78 (canonical=1) and
 * PredictionSchemeNormalOctahedronDecodingTransform.cs:47-67 (canonical=0); asserts dropped and
 * AddAsUnsigned = plain mod-2^32 add (B-9). */
static inline void oct_original(const octbox_t *t, int canonical, const int32_t *pred_in, const int32_t *corr,
                                int32_t *out) {
  int32_t p0 = (int32_t)((uint32_t)pred_in[0] - (uint32_t)t->center);
  int32_t p1 = (int32_t)((uint32_t)pred_in[1] - (uint32_t)t->center);
  int in_diamond = oct_in_diamond(t, p0, p1);
  if (!in_diamond) oct_invert_diamond(t, &p0, &p1);
  int bottom_left = 1, rot = 0;
  if (canonical) {
    bottom_left = (p0 == 0 && p1 == 0) ? 1 : (p0 < 0 && p1 <= 0);
    rot = oct_rotation_count(p0, p1);
    if (!bottom_left) oct_rotate(&p0, &p1, rot);
  }
  int32_t o0 = oct_mod_max(t, (int32_t)((uint32_t)p0 + (uint32_t)corr[0]));
  int32_t o1 = oct_mod_max(t, (int32_t)((uint32_t)p1 + (uint32_t)corr[1]));
  if (canonical && !bottom_left) oct_rotate(&o0, &o1, (4 - rot) % 4);
  if (!in_diamond) oct_invert_diamond(t, &o0, &o1);
  out[0] = (int32_t)((uint32_t)o0 + (uint32_t)t->center);
  out[1] = (int32_t)((uint32_t)o1 + (uint32_t)t->center);
}
/* Delta decoder (PredictionSchemeDeltaDecoder.cs:23-37) with an octahedron transform; nc = 2.
 * max_q -> tool box per PredictionSchemeNormalOctahedronTransform.cs:44-53. */
void orc_delta_oct(const int32_t *corr, uint32_t n, int32_t max_q, int canonical, int32_t *out) {
  octbox_t t;
  int msb = -1;
  for (uint32_t v = (uint32_t)max_q; v; v >>= 1) ++msb;
  octbox_set(&t, msb + 1);
  int32_t zero[2] = {0, 0};
  if (n == 0) return;
  oct_original(&t, canonical, zero, corr, out);
  for (uint32_t i = 1; i < n; ++i) oct_original(&t, canonical, out + 2 * (uint64_t)(i - 1), corr + 2 * (uint64_t)i, out + 2 * (uint64_t)i);
}

/* OctahedronToolBox.CanonicalizeIntegerVector (:121-137), 64-bit products as the bitstream defines them (the C#
 * multiplies in int, which overflows for 30-bit centres). */
static void oct_canonicalize_vector(const octbox_t *t, int32_t v[3]) {
  const int64_t abs_sum = (int64_t)iabs32(v[0]) + (int64_t)iabs32(v[1]) + (int64_t)iabs32(v[2]);
  if (abs_sum == 0) {
    v[0] = t->center;
  } else {
    v[0] = (int32_t)(((int64_t)v[0] * (int64_t)t->center) / abs_sum);
    v[1] = (int32_t)(((int64_t)v[1] * (int64_t)t->center) / abs_sum);
    const int32_t rest = t->center - iabs32(v[0]) - iabs32(v[1]);
    v[2] = v[2] >= 0 ? rest : -rest;
  }
}
/* OctahedronToolBox.IntegerVectorToQuantizedOctahedralCoords (:61-77) + CanonicalizeOctahedralCoords (:28-54) */
static void oct_vector_to_coords(const octbox_t *t, const int32_t v[3], int32_t *os, int32_t *ot) {
  int32_t s, tt;
  if (v[0] >= 0) {
    s = v[1] + t->center;
    tt = v[2] + t->center;
  } else {
    s = v[1] < 0 ? iabs32(v[2]) : t->max_value - iabs32(v[2]);
    tt = v[2] < 0 ? iabs32(v[1]) : t->max_value - iabs32(v[1]);
  }
  const int32_t mv = t->max_value, ce = t->center;
  if ((s == 0 && tt == 0) || (s == 0 && tt == mv) || (s == mv && tt == 0)) { s = mv; tt = mv; }
  else if (s == 0 && tt > ce) tt = ce - (tt - ce);
  else if (s == mv && tt < ce) tt = ce + (ce - tt);
  else if (tt == mv && s < ce) s = ce + (ce - s);
  else if (tt == 0 && s > ce) s = ce - (s - ce);
  *os = s;
  *ot = tt;
}
/* Vector<long>.AbsSum (D/IO/Core/Vector.cs:211-226) with the absolute values the C# forgets: saturates at INT64_MAX */
static int64_t abs_sum3_sat(const int64_t v[3]) {
  int64_t r = 0;
  for (int i = 0; i < 3; ++i) {
    if (v[i] == INT64_MIN) return INT64_MAX;
    const int64_t a = v[i] < 0 ? -v[i] : v[i];
    if (r > INT64_MAX - a) return INT64_MAX;
    r += a;
  }
  return r;
}

/* MeshPredictionSchemeGeometricNormalDecoder.ComputeOriginalValues (:44-69) over
 * MeshPredictionSchemeGeometricNormalPredictorArea.ComputePredictedValue (:11-60, TriangleArea mode) and the octahedron
 * transform, in the bitstream's semantics where the C# is defective (the corner iterator skips
 * its first corner, AbsSum takes no absolute values, the result is returned as (n0, n1, n0)).  The predicted normal of
 * an entry is the sum of the cross products of the triangles around the entry's vertex in POSITION space -- no
 * dependence on other normals: the scheme is point-parallel -- scaled below 2^29, canonicalised to |n|_1 = centre,
 * negated when the entry's rABS-coded flip bit is set, and mapped to octahedral coordinates.
 * m: maps of the normal attribute's decoder; pm / pos_q / n_pos: maps and decoded portable values of the position
 * attribute (parent).  The position of a corner is the position entry of that corner's vertex in pm (the same
 * lookup GetPositionForCorner :27-32 performs through the entry-to-point map). */
int orc_geometric_normal_oct(const int32_t *corr, uint32_t n, int32_t max_q, int canonical, const orc_mesh_maps *m,
                             const int32_t *pos_q, uint32_t n_pos, const orc_mesh_maps *pm, const uint8_t *flip,
                             int32_t *out) {
  octbox_t t;
  int msb = -1;
  for (uint32_t v = (uint32_t)max_q; v; v >>= 1) ++msb;
  octbox_set(&t, msb + 1);
  if (n == 0) return ORC_OK;
  if (!m || !pm || m->n_entries < n) return ORC_ERR_MAPS;
  int bad = 0;
#define POS_OF(corner, dst)                                                                   \
  do {                                                                                        \
    uint32_t c_ = (corner);                                                                   \
    if (c_ >= pm->n_corners) return ORC_ERR_MAPS;                                             \
    uint32_t v_ = pm->corner_to_vertex[c_];                                                   \
    if (v_ >= pm->n_vertices) return ORC_ERR_MAPS;                                            \
    int32_t e_ = pm->vertex_to_data[v_];                                                      \
    if (e_ < 0 || (uint32_t)e_ >= n_pos) return ORC_ERR_MAPS;                                 \
    for (int k_ = 0; k_ < 3; ++k_) (dst)[k_] = pos_q[3ull * (uint32_t)e_ + k_];               \
  } while (0)
  for (uint32_t p = 0; p < n; ++p) {
    const uint32_t start = m->data_to_corner[p];
    if (start >= m->n_corners) return ORC_ERR_MAPS;
    int64_t cent[3], nx[3], pv[3];
    uint64_t nrm[3] = {0, 0, 0};
    POS_OF(start, cent);
    uint32_t corner = start;
    int left = 1;
    uint64_t guard = 0;
    while (corner != 0xFFFFFFFFu) { /* VertexCornersIterator (D/IO/Mesh/VertexCornersIterator.cs:19-44), start included */
      if (++guard > (uint64_t)m->n_corners + 2) return ORC_ERR_MAPS;
      POS_OF(m_next(corner), nx);
      POS_OF(m_prev(corner), pv);
      int64_t dn[3], dp[3];
      for (int k = 0; k < 3; ++k) { dn[k] = nx[k] - cent[k]; dp[k] = pv[k] - cent[k]; }
      nrm[0] += (uint64_t)dn[1] * (uint64_t)dp[2] - (uint64_t)dn[2] * (uint64_t)dp[1]; /* CrossProduct, summed as unsigned :33-36 */
      nrm[1] += (uint64_t)dn[2] * (uint64_t)dp[0] - (uint64_t)dn[0] * (uint64_t)dp[2];
      nrm[2] += (uint64_t)dn[0] * (uint64_t)dp[1] - (uint64_t)dn[1] * (uint64_t)dp[0];
      if (left) {
        corner = m_next(m_opp(m, m_next(corner), &bad)); /* SwingLeft */
        if (bad) return ORC_ERR_MAPS;
        if (corner == 0xFFFFFFFFu) {
          corner = m_prev(m_opp(m, m_prev(start), &bad)); /* SwingRight(start) */
          left = 0;
        } else if (corner == start) {
          corner = 0xFFFFFFFFu;
        }
      } else {
        corner = m_prev(m_opp(m, m_prev(corner), &bad));
      }
      if (bad) return ORC_ERR_MAPS;
    }
    int64_t nv[3] = {(int64_t)nrm[0], (int64_t)nrm[1], (int64_t)nrm[2]};
    const int64_t upper = 1ll << 29; /* :38 */
    const int64_t abs_sum = abs_sum3_sat(nv);
    if (abs_sum > upper) {           /* :49-53 */
      const int64_t q = abs_sum / upper;
      for (int k = 0; k < 3; ++k) nv[k] /= q;
    }
    int32_t v3[3] = {(int32_t)nv[0], (int32_t)nv[1], (int32_t)nv[2]};
    oct_canonicalize_vector(&t, v3); /* ...GeometricNormalDecoder.cs:55 */
    if (flip[p])                     /* :58-61 */
      for (int k = 0; k < 3; ++k) v3[k] = (int32_t)(0u - (uint32_t)v3[k]);
    int32_t pred[2];
    oct_vector_to_coords(&t, v3, &pred[0], &pred[1]);
    oct_original(&t, canonical, pred, corr + 2ull * p, out + 2ull * p); /* :65 */
  }
#undef POS_OF
  return ORC_OK;
}

/* D/IO/Attributes/AttributeQuantizationTransform.cs:179-199 + D/IO/Core/Dequantizer.cs:14-23.
 * delta = range / (float)max_q : one rounded binary32 division;
 * out = (float)q * delta + min : two separately rounded binary32 operations. */
void orc_dequantize(const int32_t *q, uint32_t n, int nc, const float *mn, float range, int bits, float *out) {
  int32_t max_q = (int32_t)((1u << bits) - 1u);
  volatile float delta = range / (float)max_q;
  for (uint64_t i = 0; i < (uint64_t)n; ++i)
    for (int c = 0; c < nc; ++c) {
      volatile float prod = (float)q[i * nc + c] * delta;
      out[i * nc + c] = prod + mn[c];
    }
}

/* D/IO/Attributes/AttributeOctahedronTransform.cs:82-102 + OctahedronToolBox.cs:139-142,220-239
 * with B-10 / B-11 applied: Float32 target, norm = x^2+y^2+z^2.  Arithmetic types follow the C#:
 * x,y,z and the squared norm are binary32; 1/sqrt and the final products are binary64, then
 * rounded to binary32. */
void orc_oct_to_unit(const int32_t *st, uint32_t n, int bits, float *out) {
  octbox_t t;
  octbox_set(&t, bits);
  volatile float scale = 2.0f / (float)t.max_value; /* :19 */
  for (uint64_t i = 0; i < (uint64_t)n; ++i) {
    volatile float ys = (float)st[2 * i] * scale;
    volatile float zs = (float)st[2 * i + 1] * scale;
    float y = ys - 1.0f, z = zs - 1.0f; /* :141 */
    volatile float x1 = 1.0f - fabsf(y);
    float x = x1 - fabsf(z); /* :224 */
    float x_offset = (-x < 0) ? 0.0f : -x;
    y += (y < 0) ? x_offset : -x_offset;
    z += (z < 0) ? x_offset : -x_offset;
    volatile float xx = x * x, yy = y * y, zz = z * z;
    volatile float s1 = xx + yy;
    float norm_squared = s1 + zz; /* B-11 */
    if ((double)norm_squared < 1E-6) {
      out[3 * i] = 0;
      out[3 * i + 1] = 0;
      out[3 * i + 2] = 0;
    } else {
      double d = 1.0 / sqrt((double)norm_squared); /* 1.0f / Math.Sqrt(double) */
      out[3 * i] = (float)((double)x * d);
      out[3 * i + 1] = (float)((double)y * d);
      out[3 * i + 2] = (float)((double)z * d);
    }
  }
}

static int dtype_len(int dt) { /* D/IO/Constants.cs:133-143 */
  switch (dt) {
    case 1: case 2: case 11: return 1;
    case 3: case 4: return 2;
    case 5: case 6: case 9: return 4;
    case 7: case 8: case 10: return 8;
    default: return -1;
  }
}

/* D/IO/Attributes/SequentialIntegerAttributeDecoder.cs:103-160 with B-12 applied (tight stride):
 * keep the low sizeof(T) bytes of every int32. */
void orc_narrow(const int32_t *q, uint64_t count, int data_type, uint8_t *out) {
  int sz = dtype_len(data_type);
  for (uint64_t i = 0; i < count; ++i) {
    uint32_t u = (uint32_t)q[i];
    for (int b = 0; b < sz; ++b) out[i * sz + b] = (uint8_t)(u >> (8 * b));
  }
}

uint64_t orc_fnv1a(const uint8_t *p, uint64_t n, uint64_t h) {
  if (h == 0) h = 1469598103934665603ull;
  for (uint64_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

/* ------------------------------------------------------------------------- */
/* container walk                                                             */
/* ------------------------------------------------------------------------- */

/* D/IO/Metadata/MetadataDecoder.cs:5-49 (skip only; the sub-metadata crash at :45 is not mirrored) */
static void skip_metadata_element(rd_t *r, int depth) {
  if (depth > 64) {
    r->err = ORC_ERR_UNSUPPORTED;
    return;
  }
  uint64_t n = rd_varint(r);
  for (uint64_t i = 0; i < n && !r->err; ++i) {
    uint8_t ks = rd_u8(r);
    if (rd_need(r, ks)) r->pos += ks;
    uint8_t vs = rd_u8(r);
    if (rd_need(r, vs)) r->pos += vs;
  }
  uint64_t ns = rd_varint(r);
  for (uint64_t i = 0; i < ns && !r->err; ++i) {
    uint8_t ks = rd_u8(r);
    if (rd_need(r, ks)) r->pos += ks;
    skip_metadata_element(r, depth + 1);
  }
}
static void skip_metadata(rd_t *r) {
  uint64_t n = rd_varint(r);
  for (uint64_t i = 0; i < n && !r->err; ++i) {
    (void)rd_varint(r);
    skip_metadata_element(r, 0);
  }
  skip_metadata_element(r, 0);
}

/* implemented in orc_eb.c: Edgebreaker connectivity + per-decoder traversal maps.  Returns
 * ORC_ERR_UNSUPPORTED when that file is not linked in. */
int orc_eb_decode_connectivity(const uint8_t *buf, uint64_t len, uint64_t *pos, int traversal_type, orc_result *res);
int orc_eb_build_maps(orc_result *res, const uint8_t *dec_ids, int n_dec);
int orc_seq_mesh_connectivity(const uint8_t *buf, uint64_t len, uint64_t *pos, orc_result *res);
int orc_rabs_bits(const uint8_t *buf, uint64_t len, uint64_t *pos, uint32_t n_bits, uint8_t *bits_out);

/* PORTABLE(int-like): D/IO/Attributes/SequentialIntegerAttributeDecoder.cs:23-101 */
static int decode_portable(rd_t *r, orc_result *res, orc_attr *a, const orc_mesh_maps *maps, int n_maps) {
  uint32_t n = a->n_entries;
  if (a->seq_type == 0) { /* generic: SequentialAttributeDecoder.cs:75-86 */
    uint64_t stride = (uint64_t)dtype_len(a->data_type) * a->nc;
    a->out_bytes = stride * n;
    if (!rd_need(r, a->out_bytes)) return r->err;
    a->out = (uint8_t *)malloc(a->out_bytes ? a->out_bytes : 1);
    memcpy(a->out, r->p + r->pos, a->out_bytes);
    r->pos += a->out_bytes;
    return ORC_OK;
  }
  a->pred_method = rd_i8(r); /* :25 */
  if (r->err) return r->err;
  if (a->pred_method < -2 || a->pred_method >= 7) return ORC_ERR_PRED; /* :26 */
  a->transform = -1;
  if (a->pred_method != -2) {
    a->transform = rd_i8(r); /* :30 */
    if (r->err) return r->err;
    if (a->transform < -1 || a->transform >= 4) return ORC_ERR_PRED; /* :31 */
  }
  /* which scheme objects exist: :46-51 (Wrap only) / SequentialNormalAttributeDecoder.cs:19-27 (B-8) */
  int has_scheme = 0;
  if (a->pred_method != -2) {
    if (a->seq_type == 3)
      has_scheme = (a->transform == 2 || a->transform == 3);
    else
      has_scheme = (a->transform == 1);
  }
  int is_mesh = res->geom_type == 1;
  int mesh_scheme = 0; /* PredictionSchemeDecoderFactory.cs:9-75 */
  if (has_scheme && is_mesh && res->method == 1) {
    /* Edgebreaker meshes have corner table + encoding data: mesh schemes apply */
    if (a->pred_method == 1)
      mesh_scheme = 1;
    else if (a->pred_method == 5 && a->nc_portable == 2 && a->transform == 1)
      mesh_scheme = 5; /* TexCoordsPortable */
    else if (a->pred_method == 4 && a->transform == 1)
      mesh_scheme = 4; /* ConstrainedMultiParallelogram */
    else if (a->pred_method == 6 && a->seq_type == 3)
      mesh_scheme = 6; /* GeometricNormal */
    else if (a->pred_method != 0)
      return ORC_ERR_UNSUPPORTED; /* multi-parallelogram, deprecated texcoords, geometric normal: SURVEY 8f-3 */
  }
  int ncp = a->nc_portable;
  uint64_t nv = (uint64_t)n * ncp;
  a->compressed = rd_u8(r); /* :61 */
  if (r->err) return r->err;
  a->symbols = (uint32_t *)calloc(nv ? nv : 1, 4);
  a->corr = (int32_t *)calloc(nv ? nv : 1, 4);
  a->qints = (int32_t *)calloc(nv ? nv : 1, 4);
  a->scheme = -1;
  if (a->compressed > 0) {
    int st = orc_decode_symbols(r->p, r->len, &r->pos, (uint32_t)nv, (uint32_t)ncp, a->symbols, a); /* :65 */
    if (st) return st;
  } else { /* :68-84 with B-6 applied */
    uint8_t nb = rd_u8(r);
    if (r->err) return r->err;
    if (nb > 4) return ORC_ERR_UNSUPPORTED;
    if (!rd_need(r, (uint64_t)nb * nv)) return r->err;
    for (uint64_t i = 0; i < nv; ++i) {
      uint32_t v = 0;
      for (int b = 0; b < nb; ++b) v |= (uint32_t)r->p[r->pos++] << (8 * b);
      a->symbols[i] = v;
    }
  }
  /* zig-zag iff no scheme or the transform's corrections may be negative: :86-90 with B-5 */
  int zigzag = (!has_scheme) || (a->transform == 1);
  for (uint64_t i = 0; i < nv; ++i) a->corr[i] = zigzag ? orc_zigzag(a->symbols[i]) : (int32_t)a->symbols[i];
  if (!has_scheme) {
    memcpy(a->qints, a->corr, nv * 4);
    return ORC_OK;
  }
  /* PRED_DATA :93 */
  uint8_t *orient = NULL;
  uint32_t n_or = 0;
  if (mesh_scheme == 5) { /* MeshPredictionSchemeTexCoordsPortableDecoder.DecodePredictionData :68-84, before the transform data */
    int32_t no = rd_i32(r);
    if (r->err) return r->err;
    if (no < 0 || (uint64_t)no > nv + 1) return ORC_ERR_PRED; /* at most one flag per entry is ever consumed */
    n_or = (uint32_t)no;
    orient = (uint8_t *)malloc(n_or ? n_or : 1);
    int st = orc_rabs_bits(r->p, r->len, &r->pos, n_or, orient);
    if (st) { free(orient); return st; }
    int last = 1;
    for (uint32_t i = 0; i < n_or; ++i) { /* :76-82 */
      if (orient[i] == 0) last = !last;
      orient[i] = (uint8_t)last;
    }
  }
  uint8_t *crease[4] = {NULL, NULL, NULL, NULL};
  uint32_t n_crease[4] = {0, 0, 0, 0};
  if (mesh_scheme == 4) { /* ...ConstrainedMultiParallelogramDecoder.DecodeTransformData :119-141 (v2.2: no mode byte) */
    int st = ORC_OK;
    for (int i = 0; i < 4 && !st; ++i) {
      uint64_t nf = rd_varint(r);
      if (r->err) { st = r->err; break; }
      if (nf > 4ull * (uint64_t)n + 4ull) { st = ORC_ERR_PRED; break; } /* an entry consumes at most four flags */
      n_crease[i] = (uint32_t)nf;
      if (nf > 0) {
        crease[i] = (uint8_t *)malloc((size_t)nf);
        st = orc_rabs_bits(r->p, r->len, &r->pos, n_crease[i], crease[i]);
      }
    }
    if (st) { for (int i = 0; i < 4; ++i) free(crease[i]); return st; }
  }
  if (a->transform == 1) { /* PredictionSchemeWrapDecodingTransform.cs:69-75 */
    a->xf_a = rd_i32(r);
    a->xf_b = rd_i32(r);
    int64_t diff = (int64_t)a->xf_b - (int64_t)a->xf_a; /* WrapTransform.cs:90-91 (int overflow -> negative) */
    if (r->err || a->xf_a > a->xf_b || (int32_t)diff < 0 || diff >= 2147483647ll) {
      free(orient);
      for (int i = 0; i < 4; ++i) free(crease[i]);
      return r->err ? r->err : ORC_ERR_WRAP;
    }
    if (nv > 0) {
      if (mesh_scheme == 5) {
        if (!maps || a->decoder_id >= n_maps) { free(orient); return ORC_ERR_MAPS; }
        /* parent attribute: PointCloud.GetNamedAttributeId(Position) = the FIRST position attribute, in its portable
         * form (SequentialAttributeDecoder.InitPredictionScheme :58-72); it must have been decoded already */
        const orc_attr *pos = NULL;
        for (int i = 0; i < res->n_attrs && &res->attrs[i] != a && !pos; ++i)
          if (res->attrs[i].att_type == 0) pos = &res->attrs[i];
        if (pos && (pos->nc_portable != 3 || !pos->qints)) pos = NULL;
        if (!pos || pos->decoder_id >= n_maps) { free(orient); return ORC_ERR_PRED; }
        int st = orc_texcoords_portable_wrap(a->corr, n, a->xf_a, a->xf_b, &maps[a->decoder_id], pos->qints, pos->n_entries,
                                             &maps[pos->decoder_id], orient, n_or, a->qints);
        free(orient);
        orient = NULL;
        if (st) return st;
      } else if (mesh_scheme == 4) {
        int st = (!maps || a->decoder_id >= n_maps) ? ORC_ERR_MAPS
                 : orc_cmp_wrap(a->corr, n, ncp, a->xf_a, a->xf_b, &maps[a->decoder_id], crease, n_crease, a->qints);
        for (int i = 0; i < 4; ++i) { free(crease[i]); crease[i] = NULL; }
        if (st) return st;
      } else if (mesh_scheme) {
        if (!maps || a->decoder_id >= n_maps) return ORC_ERR_MAPS;
        int st = orc_parallelogram_wrap(a->corr, n, ncp, a->xf_a, a->xf_b, &maps[a->decoder_id], a->qints);
        if (st) return st;
      } else {
        orc_delta_wrap(a->corr, n, ncp, a->xf_a, a->xf_b, a->qints);
      }
    }
    free(orient);
    for (int i = 0; i < 4; ++i) free(crease[i]);
  } else { /* octahedron transforms */
    a->xf_a = rd_i32(r); /* max_quantized_value */
    if (a->transform == 3) a->xf_b = rd_i32(r); /* center_value (ignored) ...CanonicalizedDecodingTransform.cs:80-84 */
    if (r->err) return r->err;
    if (a->xf_a % 2 == 0) return ORC_ERR_QUANT; /* ...OctahedronTransform.cs:50 */
    int msb = -1;
    for (uint32_t v = (uint32_t)a->xf_a; v; v >>= 1) ++msb;
    if (msb + 1 < 2 || msb + 1 > 30) return ORC_ERR_QUANT; /* OctahedronToolBox.cs:15 */
    if (mesh_scheme == 6) { /* MeshPredictionSchemeGeometricNormalDecoder.DecodePredictionData :72-82: transform data, then the flip bits */
      uint8_t *flip = (uint8_t *)malloc(n ? n : 1);
      int st = orc_rabs_bits(r->p, r->len, &r->pos, n, flip);
      if (!st && nv > 0) {
        const orc_attr *pos = NULL;
        for (int i = 0; i < res->n_attrs && &res->attrs[i] != a && !pos; ++i)
          if (res->attrs[i].att_type == 0) pos = &res->attrs[i];
        if (pos && (pos->nc_portable != 3 || !pos->qints)) pos = NULL;
        if (!maps || a->decoder_id >= n_maps) st = ORC_ERR_MAPS;
        else if (!pos || pos->decoder_id >= n_maps) st = ORC_ERR_PRED;
        else st = orc_geometric_normal_oct(a->corr, n, a->xf_a, a->transform == 3, &maps[a->decoder_id], pos->qints,
                                           pos->n_entries, &maps[pos->decoder_id], flip, a->qints);
      }
      free(flip);
      return st;
    }
    if (mesh_scheme) return ORC_ERR_UNSUPPORTED;           /* parallelogram on normals: not a Draco combination */
    if (nv > 0) orc_delta_oct(a->corr, n, a->xf_a, a->transform == 3, a->qints);
  }
  return ORC_OK;
}

/* XFORM_PARAMS + store: SequentialQuantizationAttributeDecoder.cs:26-47,
 * SequentialNormalAttributeDecoder.cs:38-50 (B-7), SequentialIntegerAttributeDecoder.cs:14-21,103-160 */
static int decode_xform_params(rd_t *r, orc_attr *a) {
  if (a->seq_type == 2) { /* AttributeQuantizationTransform.cs:110-121 */
    for (int c = 0; c < a->nc; ++c) {
      float f = rd_f32(r);
      if (c < 4) a->qmin[c] = f;
    }
    a->qrange = rd_f32(r);
    a->qbits = rd_u8(r);
    if (r->err) return r->err;
    if (a->qbits < 1 || a->qbits > 30) return ORC_ERR_QUANT;
  } else if (a->seq_type == 3) { /* AttributeOctahedronTransform.cs:39-42 */
    a->qbits = rd_u8(r);
    if (r->err) return r->err;
    if (a->qbits < 2 || a->qbits > 30) return ORC_ERR_QUANT;
  }
  return ORC_OK;
}
static int store_values(orc_attr *a) {
  uint32_t n = a->n_entries;
  if (a->seq_type == 0) return ORC_OK;
  if (a->seq_type == 2) {
    a->out_bytes = (uint64_t)n * a->nc * 4;
    a->out = (uint8_t *)malloc(a->out_bytes ? a->out_bytes : 1);
    orc_dequantize(a->qints, n, a->nc, a->qmin, a->qrange, a->qbits, (float *)a->out);
  } else if (a->seq_type == 3) {
    a->out_bytes = (uint64_t)n * 12;
    a->out = (uint8_t *)malloc(a->out_bytes ? a->out_bytes : 1);
    orc_oct_to_unit(a->qints, n, a->qbits, (float *)a->out);
  } else {
    if (a->data_type < 1 || a->data_type > 6) return ORC_ERR_UNSUPPORTED; /* :137-138 */
    int sz = dtype_len(a->data_type);
    a->out_bytes = (uint64_t)n * a->nc * sz;
    a->out = (uint8_t *)malloc(a->out_bytes ? a->out_bytes : 1);
    orc_narrow(a->qints, (uint64_t)n * a->nc, a->data_type, a->out);
  }
  return ORC_OK;
}

static int decode_impl(const uint8_t *buf, uint64_t len, const orc_mesh_maps *maps_in, int n_maps_in,
                       uint64_t attr_off_in, uint32_t n_points_in, orc_result *res) {
  rd_t r = {buf, len, 0, 0};
  /* header: D/IO/DracoDecoder.cs:44-64 */
  if (!rd_need(&r, 5)) return ORC_ERR_EOF;
  if (memcmp(buf, "DRACO", 5) != 0) return ORC_ERR_MAGIC;
  r.pos = 5;
  res->ver_major = rd_u8(&r);
  res->ver_minor = rd_u8(&r);
  res->geom_type = rd_u8(&r);
  res->method = rd_u8(&r);
  res->flags = rd_u16(&r);
  if (r.err) return r.err;
  if (res->ver_major != 2 || res->ver_minor != 2) return ORC_ERR_UNSUPPORTED; /* build targets v2.2 (Appendix A) */
  if (res->flags & 0x8000) skip_metadata(&r);                                  /* :26-29 */
  if (r.err) return r.err;
  const orc_mesh_maps *maps = maps_in;
  int n_maps = n_maps_in;
  int eb = 0;
  if (res->geom_type == 0) { /* point cloud: B-1, upstream sequential container */
    if (res->method != 0) return ORC_ERR_UNSUPPORTED; /* kd-tree coding: absent from the reference */
    int32_t np = rd_i32(&r);
    if (r.err) return r.err;
    if (np < 0) return ORC_ERR_ATTR;
    res->n_points = (uint32_t)np;
  } else if (res->geom_type == 1) {
    if (res->method == 0) {
      int st = orc_seq_mesh_connectivity(buf, len, &r.pos, res);
      if (st) return st;
    } else if (res->method == 1) {
      if (attr_off_in) { /* the caller decoded connectivity (as the C# host does for the GPU path) */
        if (attr_off_in > len || !maps_in) return ORC_ERR_MAPS;
        r.pos = attr_off_in;
        res->n_points = n_points_in;
      } else {
        uint8_t tt = rd_u8(&r); /* DracoDecoder.cs:80 */
        if (r.err) return r.err;
        int st = orc_eb_decode_connectivity(buf, len, &r.pos, tt, res);
        if (st) return st;
      }
      eb = 1;
    } else
      return ORC_ERR_UNSUPPORTED;
  } else
    return ORC_ERR_UNSUPPORTED;

  /* ATTRIBUTES: D/IO/ConnectivityDecoder.cs:16-44 */
  res->attr_section_off = r.pos;
  int n_dec = rd_u8(&r);
  if (r.err) return r.err;
  res->n_decoders = n_dec;
  uint8_t *ids = (uint8_t *)calloc((size_t)(n_dec ? n_dec : 1), 3);
  int status = ORC_OK;
  if (eb) { /* DEC_ID: MeshEdgeBreakerDecoder.cs:642-662 */
    for (int i = 0; i < n_dec; ++i) {
      ids[3 * i] = rd_u8(&r);
      ids[3 * i + 1] = rd_u8(&r);
      ids[3 * i + 2] = rd_u8(&r);
    }
    if (r.err) status = r.err;
    if (!status && !attr_off_in) {
      status = orc_eb_build_maps(res, ids, n_dec);
      maps = res->maps;
      n_maps = res->n_maps;
    }
  }
  /* DEC_DATA: AttributesDecoder.cs:19-63 + SequentialAttributeDecodersController.cs:16-27 */
  int cap = 0;
  int *dec_first = (int *)calloc((size_t)(n_dec ? n_dec : 1), sizeof(int));
  int *dec_count = (int *)calloc((size_t)(n_dec ? n_dec : 1), sizeof(int));
  for (int d = 0; d < n_dec && !status; ++d) {
    uint64_t na = rd_varint(&r);
    if (r.err) { status = r.err; break; }
    if (na > (len - r.pos)) { status = ORC_ERR_EOF; break; }
    dec_first[d] = res->n_attrs;
    dec_count[d] = (int)na;
    if (res->n_attrs + (int)na > cap) {
      cap = (res->n_attrs + (int)na) * 2;
      res->attrs = (orc_attr *)realloc(res->attrs, (size_t)cap * sizeof(orc_attr));
    }
    for (uint64_t i = 0; i < na; ++i) {
      orc_attr *a = &res->attrs[res->n_attrs];
      memset(a, 0, sizeof *a);
      a->decoder_id = d;
      a->scheme = -1;
      a->pred_method = -2;
      a->transform = -1;
      a->att_type = rd_u8(&r);
      a->data_type = rd_u8(&r);
      a->nc = rd_u8(&r);
      a->normalized = rd_u8(&r) != 0;
      a->unique_id = (uint32_t)rd_varint(&r);
      res->n_attrs++;
      if (r.err) { status = r.err; break; }
      if (a->att_type >= 5 || a->data_type == 0 || a->data_type >= 12 || a->nc == 0) { status = ORC_ERR_ATTR; break; }
    }
    for (int i = 0; i < (int)na && !status; ++i) {
      orc_attr *a = &res->attrs[dec_first[d] + i];
      a->seq_type = rd_u8(&r);
      if (r.err) { status = r.err; break; }
      if (a->seq_type > 3) { status = ORC_ERR_UNSUPPORTED; break; } /* controller :78 */
      if (a->seq_type == 2 && a->data_type != 9) status = ORC_ERR_ATTR; /* SequentialQuantizationAttributeDecoder.cs:13 */
      if (a->seq_type == 3 && (a->data_type != 9 || a->nc != 3)) status = ORC_ERR_ATTR; /* Normal :14-15 */
      a->nc_portable = (a->seq_type == 3) ? 2 : a->nc; /* AttributeOctahedronTransform.cs:23-26, B-8 */
      if (a->nc > 4 && a->seq_type == 2) status = ORC_ERR_UNSUPPORTED; /* q-min array is float[4] here */
    }
  }
  /* DEC_PAYLOAD: AttributesDecoder.cs:65-70 -- all PORTABLE of a decoder, then all XFORM_PARAMS */
  for (int d = 0; d < n_dec && !status; ++d) {
    uint32_t n_entries = res->n_points; /* LinearSequencer.cs:7-13 (B-2) */
    if (eb) {
      if (!maps || d >= n_maps) { status = ORC_ERR_MAPS; break; }
      n_entries = (uint32_t)maps[d].n_entries; /* MeshAttributeIndicesEncodingObserver.cs:14-21 */
    }
    for (int i = 0; i < dec_count[d] && !status; ++i) {
      orc_attr *a = &res->attrs[dec_first[d] + i];
      a->n_entries = n_entries;
      status = decode_portable(&r, res, a, maps, n_maps);
    }
    for (int i = 0; i < dec_count[d] && !status; ++i) status = decode_xform_params(&r, &res->attrs[dec_first[d] + i]);
    for (int i = 0; i < dec_count[d] && !status; ++i) status = store_values(&res->attrs[dec_first[d] + i]);
  }
  res->end_off = r.pos;
  free(ids);
  free(dec_first);
  free(dec_count);
  return status;
}

int orc_decode(const uint8_t *buf, uint64_t len, const orc_mesh_maps *maps, int n_maps, orc_result **out) {
  return orc_decode_ex(buf, len, maps, n_maps, 0, 0, out);
}

/* attr_section_off != 0: mesh connectivity was decoded by the caller, who passes the per-decoder maps,
 * the offset of ATTRIBUTES and the point count (mirrors dcb_set_attr_section / dcb_set_mesh_maps). */
int orc_decode_ex(const uint8_t *buf, uint64_t len, const orc_mesh_maps *maps, int n_maps, uint64_t attr_section_off,
                  uint32_t n_points, orc_result **out) {
  orc_result *res = (orc_result *)calloc(1, sizeof *res);
  res->status = decode_impl(buf, len, maps, n_maps, attr_section_off, n_points, res);
  *out = res;
  return res->status;
}

void orc_free(orc_result *r) {
  if (!r) return;
  for (int i = 0; i < r->n_attrs; ++i) {
    free(r->attrs[i].symbols);
    free(r->attrs[i].corr);
    free(r->attrs[i].qints);
    free(r->attrs[i].out);
  }
  for (int i = 0; i < r->n_maps; ++i) {
    free((void *)r->maps[i].opposite);
    free((void *)r->maps[i].corner_to_vertex);
    free((void *)r->maps[i].data_to_corner);
    free((void *)r->maps[i].vertex_to_data);
  }
  free(r->maps);
  free(r->attrs);
  free(r->faces);
  free(r);
}

int orc_decode_bench(const uint8_t *buf, uint64_t len, uint64_t *points, uint64_t *out_bytes, uint64_t *checksum) {
  orc_result *res = NULL;
  int st = orc_decode(buf, len, NULL, 0, &res);
  if (!st) {
    if (points) *points += res->n_points;
    for (int i = 0; i < res->n_attrs; ++i) {
      if (out_bytes) *out_bytes += res->attrs[i].out_bytes;
      if (checksum && res->attrs[i].out_bytes) {
        /* cheap touch of the output so the work cannot be elided */
        const uint8_t *o = res->attrs[i].out;
        *checksum += o[0] + o[res->attrs[i].out_bytes - 1] + o[res->attrs[i].out_bytes / 2];
      }
    }
  }
  orc_free(res);
  return st;
}
