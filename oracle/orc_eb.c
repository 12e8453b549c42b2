/*
 * oracle/orc_eb.c -- CPU ORACLE (test infrastructure, NOT product): host connectivity for meshes.
 *
 * Plain-C restatement of the reference's Edgebreaker connectivity decoder and attribute traversal,
 * the producer of the parallelogram predictor's inputs.  "D/" = src/Draco/.
 *   connectivity header + symbol loop   D/IO/Mesh/MeshEdgeBreakerDecoder.cs:25-134,232-442
 *   topology split events               D/IO/Mesh/MeshEdgeBreakerDecoder.cs:136-230, IsTopologySplit :450-470
 *   attribute seams                     D/IO/Mesh/MeshEdgeBreakerDecoder.cs:502-535
 *   point ids / faces                   D/IO/Mesh/MeshEdgeBreakerDecoder.cs:537-638
 *   standard traversal symbols          D/IO/Mesh/MeshEdgeBreakerTraversalDecoder.cs:27-108
 *   valence traversal symbols           D/IO/Mesh/MeshEdgeBreakerTraversalValenceDecoder.cs:22-150
 *   rABS bit decoder                    D/IO/BitCoders/RAnsBitDecoder.cs:12-35, D/IO/Entropy/AnsDecoder.cs:12-56
 *   corner table                        D/IO/Mesh/CornerTable.cs:59-260
 *   attribute corner table              D/IO/Mesh/MeshAttributeCornerTable.cs:19-30,78-190
 *   depth-first traversal + observer    D/IO/Mesh/Traverser/DepthFirstTraverser.cs:9-99,
 *                                       MeshTraversalSequencer.cs:13-31, MeshAttributeIndicesEncodingObserver.cs:14-21
 *   decoder -> tables wiring            D/IO/Mesh/MeshEdgeBreakerDecoder.cs:640-708,710-760
 * Deviations where the C# cannot work: B-17 (rABS 1-byte init reads offset-2:
 * bitstream semantics offset-1 used), B-15 (encoding data looked up by decoder membership), the
 * VertexCornersIterator that skips its first corner (only reached without attribute data: bitstream
 * semantics used).  Predictive traversal (type 1) and MaxPredictionDegree traversal are not restated.
 *
 * Pinned by tests/test_oracle_house.py against the only upstream-produced asset of the reference
 * (house_04.obj.drc) and its source OBJ.
 */
#include "draco_oracle.h"

#include <stdlib.h>
#include <string.h>

#define INV 0xFFFFFFFFu

/* ---- byte / bit readers ---- */
typedef struct {
  const uint8_t *p;
  uint64_t len, pos;
  int err;
} rd_t;
static uint8_t e_u8(rd_t *r) {
  if (r->err || r->pos >= r->len) { r->err = ORC_ERR_EOF; return 0; }
  return r->p[r->pos++];
}
static uint64_t e_varint(rd_t *r) {
  uint64_t v = 0;
  if (r->err) return 0;
  int e = orc_varint(r->p, r->len, &r->pos, &v);
  if (e) r->err = e;
  return v;
}

/* rABS binary decoder: AnsDecoder.cs:12-56 (L = 4096, 8-bit probabilities) */
typedef struct {
  const uint8_t *buf;
  int64_t off;
  uint32_t state;
  uint8_t prob_zero;
} rabs_t;
static int rabs_start(rabs_t *a, rd_t *r) { /* RAnsBitDecoder.StartDecoding */
  a->prob_zero = e_u8(r);
  uint64_t n = e_varint(r);
  if (r->err) return r->err;
  if (r->len - r->pos < n) return ORC_ERR_EOF;
  const uint8_t *b = r->p + r->pos;
  r->pos += n;
  if (n < 1) return ORC_ERR_CONNECTIVITY;
  uint32_t x = (uint32_t)b[n - 1] >> 6;
  a->buf = b;
  if (x == 0) {
    a->off = (int64_t)n - 1;
    a->state = b[n - 1] & 0x3Fu; /* B-17 */
  } else if (x == 1) {
    if (n < 2) return ORC_ERR_CONNECTIVITY;
    a->off = (int64_t)n - 2;
    a->state = ((uint32_t)b[n - 2] | ((uint32_t)b[n - 1] << 8)) & 0x3FFFu;
  } else if (x == 2) {
    if (n < 3) return ORC_ERR_CONNECTIVITY;
    a->off = (int64_t)n - 3;
    a->state = ((uint32_t)b[n - 3] | ((uint32_t)b[n - 2] << 8) | ((uint32_t)b[n - 1] << 16)) & 0x3FFFFFu;
  } else {
    return ORC_ERR_CONNECTIVITY;
  }
  a->state += 4096u;
  if (a->state >= 4096u * 256u) return ORC_ERR_CONNECTIVITY;
  return ORC_OK;
}
static uint32_t rabs_bit(rabs_t *a) { /* AnsDecoder.RAbsRead */
  uint8_t p = (uint8_t)(256u - a->prob_zero);
  if (a->state < 4096u && a->off > 0) a->state = a->state * 256u + a->buf[--a->off];
  uint32_t x = a->state, quot = x / 256u, rem = x % 256u, xn = quot * p;
  int val = rem < p;
  a->state = val ? xn + rem : x - xn - p;
  return (uint32_t)val;
}

/* rABS-coded bit sequence at *pos (u8 prob_zero, varint size, data), as RAnsBitDecoder.StartDecoding / DecodeNextBit /
 * EndDecoding read it (D/IO/BitCoders/RAnsBitDecoder.cs:12-35).  Used by prediction schemes that carry flag streams
 * (MeshPredictionSchemeTexCoordsPortableDecoder.DecodePredictionData :68-84).  *pos ends behind the data. */
int orc_rabs_bits(const uint8_t *buf, uint64_t len, uint64_t *pos, uint32_t n_bits, uint8_t *bits_out) {
  rd_t r = {buf, len, *pos, 0};
  rabs_t a;
  int st = rabs_start(&a, &r);
  if (st) return st;
  for (uint32_t i = 0; i < n_bits; ++i) bits_out[i] = (uint8_t)rabs_bit(&a);
  *pos = r.pos;
  return ORC_OK;
}

/* ---- corner table ---- */
typedef struct {
  uint32_t *c2v, *opp;   /* [n_corners] */
  uint32_t n_corners;
  uint32_t *vcorner;     /* left-most corner per vertex */
  uint32_t n_vertices, cap_vertices;
} ctab_t;
static inline uint32_t c_next(uint32_t c) { return c == INV ? c : ((c + 1) % 3 != 0 ? c + 1 : c - 2); }
static inline uint32_t c_prev(uint32_t c) { return c == INV ? c : (c % 3 != 0 ? c - 1 : c + 2); }
/* a "view": base table or attribute table, same accessors (Opposite / Vertex / LeftMostCorner) */
typedef struct {
  const uint32_t *opp, *c2v, *vleft;
  uint32_t n_corners, n_vertices;
} view_t;
static inline uint32_t v_opp(const view_t *t, uint32_t c) { return c == INV ? c : t->opp[c]; }
static inline uint32_t v_vertex(const view_t *t, uint32_t c) { return (c == INV || c >= t->n_corners) ? c : t->c2v[c]; }
static inline uint32_t v_swing_right(const view_t *t, uint32_t c) { return c_prev(v_opp(t, c_prev(c))); }
static inline uint32_t v_swing_left(const view_t *t, uint32_t c) { return c_next(v_opp(t, c_next(c))); }
static inline uint32_t v_right_corner(const view_t *t, uint32_t c) { return c == INV ? INV : v_opp(t, c_next(c)); }
static inline uint32_t v_left_corner(const view_t *t, uint32_t c) { return c == INV ? INV : v_opp(t, c_prev(c)); }
static int v_on_boundary(const view_t *t, uint32_t v) {
  uint32_t c = t->vleft[v];
  return c == INV || v_swing_left(t, c) == INV;
}

static uint32_t ct_add_vertex(ctab_t *t) {
  if (t->n_vertices == t->cap_vertices) {
    t->cap_vertices = t->cap_vertices ? t->cap_vertices * 2 : 64;
    t->vcorner = (uint32_t *)realloc(t->vcorner, (size_t)t->cap_vertices * 4);
  }
  t->vcorner[t->n_vertices] = INV;
  return t->n_vertices++;
}
static void ct_set_opp(ctab_t *t, uint32_t a, uint32_t b) {
  t->opp[a] = b;
  t->opp[b] = a;
}

/* ---- traversal symbol sources ---- */
typedef struct {
  int type; /* 0 standard, 2 valence */
  /* standard */
  const uint8_t *sym_bits;
  uint64_t sym_len, sym_bitpos;
  /* valence */
  uint32_t *valence;
  uint32_t n_valence;
  uint32_t *ctx_syms[6];
  int64_t ctx_count[6];
  int last_symbol, active_context;
  rabs_t start_face;
  rabs_t *seams;
  int n_seams;
} trav_t;

static const uint8_t kSymbolToTopology[5] = {0, 1, 3, 5, 7}; /* C S L R E: Constants.cs:86-93 */

static uint32_t trav_symbol(trav_t *t, int *err) {
  if (t->type == 0) { /* MeshEdgeBreakerTraversalDecoder.DecodeSymbol */
    uint32_t s = orc_read_bits_lsb(t->sym_bits, t->sym_len, &t->sym_bitpos, 1, err);
    if (s == 0) return 0;
    uint32_t suffix = orc_read_bits_lsb(t->sym_bits, t->sym_len, &t->sym_bitpos, 2, err);
    return s | (suffix << 1);
  }
  if (t->active_context != -1) { /* valence: :75-99 */
    int64_t k = --t->ctx_count[t->active_context];
    if (k < 0) return 9;
    uint32_t id = t->ctx_syms[t->active_context][k];
    if (id > 4) return 9;
    t->last_symbol = kSymbolToTopology[id];
  } else {
    t->last_symbol = 7; /* the first symbol is an implied E */
  }
  return (uint32_t)t->last_symbol;
}
static void trav_new_corner(trav_t *t, const ctab_t *ct, uint32_t corner) { /* valence :100-144 */
  if (t->type != 2) return;
  uint32_t next = c_next(corner), prev = c_prev(corner);
  uint32_t vc = ct->c2v[corner], vn = ct->c2v[next], vp = ct->c2v[prev];
  switch (t->last_symbol) {
    case 0: case 1: t->valence[vn] += 1; t->valence[vp] += 1; break;
    case 5: t->valence[vc] += 1; t->valence[vn] += 1; t->valence[vp] += 2; break;
    case 3: t->valence[vc] += 1; t->valence[vn] += 2; t->valence[vp] += 1; break;
    case 7: t->valence[vc] += 2; t->valence[vn] += 2; t->valence[vp] += 2; break;
    default: break;
  }
  int av = (int)t->valence[vn];
  int cl = av < 2 ? 2 : (av > 7 ? 7 : av);
  t->active_context = cl - 2;
}

/* ---- decoder state kept between connectivity and attribute phases ---- */
typedef struct {
  ctab_t ct;
  uint32_t n_attr_data;
  uint8_t **edge_seam;     /* per attribute data: [n_corners] */
  uint8_t **vert_seam;     /* per attribute data: [n_vertices] _isVertexOnSeam */
  uint32_t **a_c2v, **a_opp, **a_vleft; /* attribute corner tables */
  uint32_t *a_nverts;
  uint8_t *is_vert_hole;
  uint32_t n_points;
} eb_state_t;

static eb_state_t *g_state_of(orc_result *res);

/* hidden pointer: orc_result has no spare field, so the state is kept in a small side table keyed by res */
#define MAX_LIVE 64
static struct { orc_result *res; eb_state_t *st; } g_live[MAX_LIVE];
static void live_put(orc_result *res, eb_state_t *st) {
  for (int i = 0; i < MAX_LIVE; ++i)
    if (!g_live[i].res) { g_live[i].res = res; g_live[i].st = st; return; }
}
static eb_state_t *g_state_of(orc_result *res) {
  for (int i = 0; i < MAX_LIVE; ++i)
    if (g_live[i].res == res) return g_live[i].st;
  return NULL;
}
static void state_free(eb_state_t *s) {
  if (!s) return;
  free(s->ct.c2v); free(s->ct.opp); free(s->ct.vcorner);
  for (uint32_t i = 0; i < s->n_attr_data; ++i) {
    if (s->edge_seam) free(s->edge_seam[i]);
    if (s->vert_seam) free(s->vert_seam[i]);
    if (s->a_c2v) free(s->a_c2v[i]);
    if (s->a_opp) free(s->a_opp[i]);
    if (s->a_vleft) free(s->a_vleft[i]);
  }
  free(s->edge_seam); free(s->vert_seam); free(s->a_c2v); free(s->a_opp); free(s->a_vleft); free(s->a_nverts); free(s->is_vert_hole);
  free(s);
}
static void live_drop(orc_result *res) {
  for (int i = 0; i < MAX_LIVE; ++i)
    if (g_live[i].res == res) { state_free(g_live[i].st); g_live[i].res = NULL; g_live[i].st = NULL; }
}

/* MeshAttributeCornerTable.RecomputeVertices(null, null): :107-155 */
static void attr_table_build(eb_state_t *s, uint32_t ai) {
  const ctab_t *ct = &s->ct;
  const uint8_t *es = s->edge_seam[ai];
  uint32_t nc = ct->n_corners;
  uint8_t *vert_seam = (uint8_t *)calloc(ct->n_vertices ? ct->n_vertices : 1, 1);
  uint32_t *aopp = (uint32_t *)malloc((size_t)(nc ? nc : 1) * 4);
  for (uint32_t c = 0; c < nc; ++c) {
    aopp[c] = es[c] ? INV : ct->opp[c];
    if (es[c]) { /* AddSeamEdge marks both end vertices of the edge opposite to c */
      vert_seam[ct->c2v[c_next(c)]] = 1;
      vert_seam[ct->c2v[c_prev(c)]] = 1;
    }
  }
  uint32_t *ac2v = (uint32_t *)malloc((size_t)(nc ? nc : 1) * 4);
  for (uint32_t c = 0; c < nc; ++c) ac2v[c] = INV;
  uint32_t cap = nc + 1, nv = 0;
  uint32_t *vleft = (uint32_t *)malloc((size_t)cap * 4);
  view_t av = {aopp, ac2v, vleft, nc, 0};
  view_t bv = {ct->opp, ct->c2v, ct->vcorner, nc, ct->n_vertices};
  for (uint32_t v = 0; v < ct->n_vertices; ++v) {
    uint32_t c = ct->vcorner[v];
    if (c == INV) continue;
    uint32_t first_vert = nv++;
    uint32_t first_c = c, act;
    if (vert_seam[v]) {
      act = v_swing_left(&av, first_c);
      while (act != INV) {
        first_c = act;
        act = v_swing_left(&av, act);
        if (act == c) break; /* C# throws: cannot happen on valid data */
      }
    }
    ac2v[first_c] = first_vert;
    vleft[first_vert] = first_c;
    act = v_swing_right(&bv, first_c);
    while (act != INV && act != first_c) {
      if (es[c_next(act)]) {
        first_vert = nv++;
        vleft[first_vert] = first_c; /* sic: the reference stores firstC here (:146) */
      }
      ac2v[act] = first_vert;
      act = v_swing_right(&bv, act);
    }
  }
  s->a_c2v[ai] = ac2v;
  s->a_opp[ai] = aopp;
  s->a_vleft[ai] = vleft;
  s->a_nverts[ai] = nv;
  s->vert_seam[ai] = vert_seam;
}

int orc_eb_decode_connectivity(const uint8_t *buf, uint64_t len, uint64_t *pos, int traversal_type, orc_result *res) {
  if (traversal_type != 0 && traversal_type != 2) return ORC_ERR_UNSUPPORTED; /* predictive: not restated */
  rd_t r = {buf, len, *pos, 0};
  eb_state_t *s = (eb_state_t *)calloc(1, sizeof *s);
  live_put(res, s);
  int status = ORC_OK;
  trav_t tv;
  memset(&tv, 0, sizeof tv);
  tv.type = traversal_type;
  tv.last_symbol = -1;
  tv.active_context = -1;
  uint32_t *stack = NULL, *split_src = NULL, *split_id = NULL, *tsplit_key = NULL, *tsplit_val = NULL, *invalid_verts = NULL;
  uint8_t *split_edge = NULL;
  uint32_t *init_corners = NULL;

#define FAIL(code) do { status = (code); goto done; } while (0)
  uint64_t n_enc_verts = e_varint(&r), n_faces = e_varint(&r);
  if (r.err) FAIL(r.err);
  if (n_faces > (1u << 28) || n_enc_verts > n_faces * 3) FAIL(ORC_ERR_CONNECTIVITY);
  if (n_enc_verts * (n_enc_verts - 1) / 2 < 3 * n_faces / 2 && n_faces > 0) FAIL(ORC_ERR_CONNECTIVITY);
  uint32_t n_attr_data = e_u8(&r);
  uint64_t n_symbols = e_varint(&r);
  if (r.err) FAIL(r.err);
  if (n_faces < n_symbols || n_faces > n_symbols + n_symbols / 3) FAIL(ORC_ERR_CONNECTIVITY);
  uint64_t n_split_symbols = e_varint(&r);
  if (r.err) FAIL(r.err);
  if (n_split_symbols > n_symbols) FAIL(ORC_ERR_CONNECTIVITY);
  s->n_attr_data = n_attr_data;
  ctab_t *ct = &s->ct;
  ct->n_corners = (uint32_t)n_faces * 3;
  ct->c2v = (uint32_t *)malloc((size_t)(ct->n_corners ? ct->n_corners : 1) * 4);
  ct->opp = (uint32_t *)malloc((size_t)(ct->n_corners ? ct->n_corners : 1) * 4);
  for (uint32_t c = 0; c < ct->n_corners; ++c) ct->c2v[c] = ct->opp[c] = INV;
  uint32_t max_verts = (uint32_t)(n_enc_verts + n_split_symbols);
  s->is_vert_hole = (uint8_t *)malloc(max_verts ? max_verts : 1);
  memset(s->is_vert_hole, 1, max_verts ? max_verts : 1);

  /* topology split events: :136-196 (v2.2: varint deltas, then one bit per event) */
  uint64_t n_splits = e_varint(&r);
  if (r.err) FAIL(r.err);
  if (n_splits > n_faces) FAIL(ORC_ERR_CONNECTIVITY);
  split_src = (uint32_t *)malloc((size_t)(n_splits ? n_splits : 1) * 4);
  split_id = (uint32_t *)malloc((size_t)(n_splits ? n_splits : 1) * 4);
  split_edge = (uint8_t *)malloc((size_t)(n_splits ? n_splits : 1));
  if (n_splits > 0) {
    uint32_t last = 0;
    for (uint64_t i = 0; i < n_splits; ++i) {
      uint32_t d = (uint32_t)e_varint(&r);
      split_src[i] = d + last;
      d = (uint32_t)e_varint(&r);
      if (r.err) FAIL(r.err);
      if (d > split_src[i]) FAIL(ORC_ERR_CONNECTIVITY);
      split_id[i] = split_src[i] - d;
      last = split_src[i];
    }
    uint64_t bp = 0;
    int be = 0;
    for (uint64_t i = 0; i < n_splits; ++i)
      split_edge[i] = (uint8_t)(orc_read_bits_lsb(buf + r.pos, len - r.pos, &bp, 1, &be) & 1u);
    if (be) FAIL(be);
    r.pos += (bp + 7) / 8;
  }
  /* Traversal_Start */
  if (traversal_type == 0) {
    uint64_t tsz = e_varint(&r);
    if (r.err) FAIL(r.err);
    if (len - r.pos < tsz) FAIL(ORC_ERR_EOF);
    tv.sym_bits = buf + r.pos;
    tv.sym_len = tsz;
    r.pos += tsz;
  }
  status = rabs_start(&tv.start_face, &r);
  if (status) goto done;
  tv.n_seams = (int)n_attr_data;
  tv.seams = (rabs_t *)calloc(n_attr_data ? n_attr_data : 1, sizeof(rabs_t));
  for (uint32_t i = 0; i < n_attr_data; ++i) {
    status = rabs_start(&tv.seams[i], &r);
    if (status) goto done;
  }
  if (traversal_type == 2) {
    tv.n_valence = max_verts;
    tv.valence = (uint32_t *)calloc(max_verts ? max_verts : 1, 4);
    for (int i = 0; i < 6; ++i) {
      uint64_t n = e_varint(&r);
      if (r.err) FAIL(r.err);
      if (n > n_faces) FAIL(ORC_ERR_CONNECTIVITY);
      if (n > 0) {
        tv.ctx_syms[i] = (uint32_t *)calloc(n, 4);
        int e = orc_decode_symbols(buf, len, &r.pos, (uint32_t)n, 1, tv.ctx_syms[i], NULL);
        if (e) FAIL(e);
        tv.ctx_count[i] = (int64_t)n;
      }
    }
  }
  /* ---- DecodeConnectivity(numSymbols): :232-442 ---- */
  {
    uint32_t sp = 0;
    stack = (uint32_t *)malloc((size_t)(n_symbols + 4) * 4);
    tsplit_key = (uint32_t *)malloc((size_t)(n_splits + 1) * 4);
    tsplit_val = (uint32_t *)malloc((size_t)(n_splits + 1) * 4);
    invalid_verts = (uint32_t *)malloc((size_t)(n_split_symbols + 1) * 4);
    uint32_t n_tsplit = 0, n_invalid = 0;
    int64_t split_top = (int64_t)n_splits - 1; /* _topologySplitData.Last() */
    const int remove_invalid = n_attr_data == 0;
    uint32_t num_faces = 0;
    int berr = 0;
    for (uint64_t sid = 0; sid < n_symbols; ++sid) {
      uint32_t face = num_faces++;
      int check_split = 0;
      uint32_t sym = trav_symbol(&tv, &berr);
      if (berr) FAIL(berr);
      uint32_t corner = 3 * face;
      if (sym == 0) { /* C */
        if (sp == 0) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t ca = stack[sp - 1];
        uint32_t vx = ct->c2v[c_next(ca)];
        if (vx >= ct->n_vertices) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t lm = ct->vcorner[vx];
        if (lm == INV) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t cb = c_next(lm);
        if (ca == cb || ct->opp[ca] != INV || ct->opp[cb] != INV) FAIL(ORC_ERR_CONNECTIVITY);
        ct_set_opp(ct, ca, corner + 1);
        ct_set_opp(ct, cb, corner + 2);
        uint32_t va_prev = ct->c2v[c_prev(ca)], vb_next = ct->c2v[c_next(cb)];
        if (vx == va_prev || vx == vb_next) FAIL(ORC_ERR_CONNECTIVITY);
        ct->c2v[corner] = vx;
        ct->c2v[corner + 1] = vb_next;
        ct->c2v[corner + 2] = va_prev;
        if (va_prev != INV) ct->vcorner[va_prev] = corner + 2;
        s->is_vert_hole[vx] = 0;
        stack[sp - 1] = corner;
      } else if (sym == 5 || sym == 3) { /* R / L */
        if (sp == 0) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t ca = stack[sp - 1];
        if (ct->opp[ca] != INV) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t oc, cl, cr;
        if (sym == 5) { oc = corner + 2; cl = corner + 1; cr = corner; }
        else { oc = corner + 1; cl = corner; cr = corner + 2; }
        ct_set_opp(ct, oc, ca);
        uint32_t nv = ct_add_vertex(ct);
        if (ct->n_vertices > max_verts) FAIL(ORC_ERR_CONNECTIVITY);
        ct->c2v[oc] = nv;
        ct->vcorner[nv] = oc;
        uint32_t vr = ct->c2v[c_prev(ca)];
        ct->c2v[cr] = vr;
        if (vr != INV) ct->vcorner[vr] = cr;
        ct->c2v[cl] = ct->c2v[c_next(ca)];
        stack[sp - 1] = corner;
        check_split = 1;
      } else if (sym == 1) { /* S */
        if (sp == 0) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t cb = stack[--sp];
        for (uint32_t k = 0; k < n_tsplit; ++k)
          if (tsplit_key[k] == (uint32_t)sid) { stack[sp++] = tsplit_val[k]; break; }
        if (sp == 0) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t ca = stack[sp - 1];
        if (ca == cb || ct->opp[ca] != INV || ct->opp[cb] != INV) FAIL(ORC_ERR_CONNECTIVITY);
        ct_set_opp(ct, ca, corner + 2);
        ct_set_opp(ct, cb, corner + 1);
        uint32_t vp = ct->c2v[c_prev(ca)];
        ct->c2v[corner] = vp;
        ct->c2v[corner + 1] = ct->c2v[c_next(ca)];
        uint32_t vb_prev = ct->c2v[c_prev(cb)];
        ct->c2v[corner + 2] = vb_prev;
        if (vb_prev != INV) ct->vcorner[vb_prev] = corner + 2;
        uint32_t cn = c_next(cb);
        uint32_t vn = ct->c2v[cn];
        if (vp >= ct->n_vertices || vn >= ct->n_vertices) FAIL(ORC_ERR_CONNECTIVITY);
        if (tv.type == 2) tv.valence[vp] += tv.valence[vn]; /* MergeVertices */
        ct->vcorner[vp] = ct->vcorner[vn];
        view_t bv = {ct->opp, ct->c2v, ct->vcorner, ct->n_corners, ct->n_vertices};
        uint32_t first = cn;
        while (cn != INV) {
          ct->c2v[cn] = vp;
          cn = v_swing_left(&bv, cn);
          if (cn == first) FAIL(ORC_ERR_CONNECTIVITY);
        }
        ct->vcorner[vn] = INV; /* MakeVertexIsolated */
        if (remove_invalid) invalid_verts[n_invalid++] = vn;
        stack[sp - 1] = corner;
      } else if (sym == 7) { /* E */
        uint32_t v0 = ct_add_vertex(ct), v1 = ct_add_vertex(ct), v2 = ct_add_vertex(ct);
        if (ct->n_vertices > max_verts) FAIL(ORC_ERR_CONNECTIVITY);
        ct->c2v[corner] = v0; ct->c2v[corner + 1] = v1; ct->c2v[corner + 2] = v2;
        ct->vcorner[v0] = corner; ct->vcorner[v1] = corner + 1; ct->vcorner[v2] = corner + 2;
        stack[sp++] = corner;
        check_split = 1;
      } else {
        FAIL(ORC_ERR_CONNECTIVITY);
      }
      trav_new_corner(&tv, ct, stack[sp - 1]);
      if (check_split) {
        uint32_t enc_sid = (uint32_t)(n_symbols - sid - 1);
        for (;;) { /* IsTopologySplit :450-470 */
          if (split_top < 0) break;
          if (split_src[split_top] > enc_sid) FAIL(ORC_ERR_CONNECTIVITY); /* encoderSplitSymbolId = -1 */
          if (split_src[split_top] != enc_sid) break;
          uint32_t edge = split_edge[split_top], enc_split = split_id[split_top];
          --split_top;
          uint32_t top = stack[sp - 1];
          uint32_t nac = edge == 1 ? c_next(top) : c_prev(top); /* RightFaceEdge = 1 */
          uint32_t dec_split = (uint32_t)(n_symbols - enc_split - 1);
          uint32_t k;
          for (k = 0; k < n_tsplit; ++k)
            if (tsplit_key[k] == dec_split) { tsplit_val[k] = nac; break; }
          if (k == n_tsplit) { tsplit_key[n_tsplit] = dec_split; tsplit_val[n_tsplit++] = nac; }
        }
      }
    }
    if (ct->n_vertices > max_verts) FAIL(ORC_ERR_CONNECTIVITY);
    init_corners = (uint32_t *)malloc((size_t)(sp + 1) * 4);
    while (sp > 0) { /* start faces: :381-418 */
      uint32_t corner = stack[--sp];
      int interior = (int)(rabs_bit(&tv.start_face) & 1u);
      if (interior) {
        if (num_faces >= n_faces) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t vn = ct->c2v[c_next(corner)];
        if (vn >= ct->n_vertices || ct->vcorner[vn] == INV) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t cb = c_next(ct->vcorner[vn]);
        uint32_t vx = ct->c2v[c_next(cb)];
        if (vx >= ct->n_vertices || ct->vcorner[vx] == INV) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t cc = c_next(ct->vcorner[vx]);
        if (corner == cb || corner == cc || cb == cc) FAIL(ORC_ERR_CONNECTIVITY);
        if (ct->opp[corner] != INV || ct->opp[cb] != INV || ct->opp[cc] != INV) FAIL(ORC_ERR_CONNECTIVITY);
        uint32_t vp = ct->c2v[c_next(cc)];
        uint32_t nc = 3 * num_faces++;
        ct_set_opp(ct, nc, corner);
        ct_set_opp(ct, nc + 1, cb);
        ct_set_opp(ct, nc + 2, cc);
        ct->c2v[nc] = vx; ct->c2v[nc + 1] = vp; ct->c2v[nc + 2] = vn;
        for (int k = 0; k < 3; ++k)
          if (ct->c2v[nc + k] < max_verts) s->is_vert_hole[ct->c2v[nc + k]] = 0;
      }
    }
    if (num_faces != n_faces) FAIL(ORC_ERR_CONNECTIVITY);
    uint32_t num_vertices = ct->n_vertices;
    view_t bv = {ct->opp, ct->c2v, ct->vcorner, ct->n_corners, ct->n_vertices};
    for (uint32_t k = 0; k < n_invalid; ++k) { /* :422-441 (bitstream semantics for the corner iterator) */
      uint32_t iv = invalid_verts[k];
      uint32_t src = num_vertices - 1;
      while (ct->vcorner[src] == INV) src = --num_vertices - 1;
      if (src < iv) continue;
      uint32_t start = ct->vcorner[src], c = start;
      int left = 1;
      while (c != INV) {
        if (ct->c2v[c] != src) FAIL(ORC_ERR_CONNECTIVITY);
        ct->c2v[c] = iv;
        if (left) {
          c = v_swing_left(&bv, c);
          if (c == INV) { c = v_swing_right(&bv, start); left = 0; }
          else if (c == start) c = INV;
        } else {
          c = v_swing_right(&bv, c);
        }
      }
      ct->vcorner[iv] = ct->vcorner[src];
      ct->vcorner[src] = INV;
      s->is_vert_hole[iv] = s->is_vert_hole[src];
      s->is_vert_hole[src] = 0;
      num_vertices--;
    }
    s->n_points = num_vertices;
  }
  /* attribute seams: DecodeAttributeConnectivitiesOnFace :502-535 */
  s->edge_seam = (uint8_t **)calloc(n_attr_data ? n_attr_data : 1, sizeof(uint8_t *));
  s->vert_seam = (uint8_t **)calloc(n_attr_data ? n_attr_data : 1, sizeof(uint8_t *));
  s->a_c2v = (uint32_t **)calloc(n_attr_data ? n_attr_data : 1, sizeof(uint32_t *));
  s->a_opp = (uint32_t **)calloc(n_attr_data ? n_attr_data : 1, sizeof(uint32_t *));
  s->a_vleft = (uint32_t **)calloc(n_attr_data ? n_attr_data : 1, sizeof(uint32_t *));
  s->a_nverts = (uint32_t *)calloc(n_attr_data ? n_attr_data : 1, 4);
  for (uint32_t i = 0; i < n_attr_data; ++i) s->edge_seam[i] = (uint8_t *)calloc(ct->n_corners ? ct->n_corners : 1, 1);
  if (n_attr_data > 0) {
    for (uint32_t ci = 0; ci < ct->n_corners; ci += 3) {
      uint32_t corners[3] = {ci, c_next(ci), c_prev(ci)};
      uint32_t src_face = ci / 3;
      for (int c = 0; c < 3; ++c) {
        uint32_t oc = ct->opp[corners[c]];
        if (oc == INV) {
          for (uint32_t i = 0; i < n_attr_data; ++i) s->edge_seam[i][corners[c]] = 2; /* boundary: seam on this side only */
          continue;
        }
        if (oc / 3 < src_face) continue;
        for (uint32_t i = 0; i < n_attr_data; ++i)
          if (rabs_bit(&tv.seams[i])) { /* AddSeamEdge marks the opposite corner as well */
            s->edge_seam[i][corners[c]] = 1;
            s->edge_seam[i][oc] = 1;
          }
      }
    }
    for (uint32_t i = 0; i < n_attr_data; ++i) {
      for (uint32_t c = 0; c < ct->n_corners; ++c) s->edge_seam[i][c] = s->edge_seam[i][c] ? 1 : 0;
      attr_table_build(s, i);
    }
  }
  /* AssignPointsToCorners: :537-638 */
  res->n_faces = (uint32_t)n_faces;
  res->faces = (uint32_t *)malloc((size_t)(ct->n_corners ? ct->n_corners : 1) * 4);
  if (n_attr_data == 0) {
    for (uint32_t c = 0; c < ct->n_corners; ++c) res->faces[c] = ct->c2v[c];
    res->n_points = s->n_points;
  } else {
    view_t bv = {ct->opp, ct->c2v, ct->vcorner, ct->n_corners, ct->n_vertices};
    uint32_t *c2p = (uint32_t *)calloc(ct->n_corners ? ct->n_corners : 1, 4);
    uint32_t n_points = 0;
    for (uint32_t v = 0; v < ct->n_vertices; ++v) {
      uint32_t c = ct->vcorner[v];
      if (c == INV) continue;
      uint32_t dedup_first = c;
      if (!s->is_vert_hole[v]) {
        for (uint32_t i = 0; i < n_attr_data; ++i) {
          const int on_seam = s->vert_seam[i][v]; /* IsCornerOnSeam(c) */
          if (!on_seam) continue;
          uint32_t vid = s->a_c2v[i][c];
          uint32_t act = v_swing_right(&bv, c);
          int found = 0;
          while (act != c) {
            if (act == INV) { free(c2p); FAIL(ORC_ERR_CONNECTIVITY); }
            if (s->a_c2v[i][act] != vid) { dedup_first = act; found = 1; break; }
            act = v_swing_right(&bv, act);
          }
          if (found) break;
        }
      }
      c = dedup_first;
      c2p[c] = n_points++;
      uint32_t prev_c = c;
      c = v_swing_right(&bv, c);
      while (c != INV && c != dedup_first) {
        int seam = 0;
        for (uint32_t i = 0; i < n_attr_data; ++i)
          if (s->a_c2v[i][c] != s->a_c2v[i][prev_c]) { seam = 1; break; }
        if (seam) c2p[c] = n_points++;
        else c2p[c] = c2p[prev_c];
        prev_c = c;
        c = v_swing_right(&bv, c);
      }
    }
    for (uint32_t c = 0; c < ct->n_corners; ++c) res->faces[c] = c2p[c];
    res->n_points = n_points;
    free(c2p);
  }
  *pos = r.pos;
done:
  free(stack); free(split_src); free(split_id); free(split_edge); free(tsplit_key); free(tsplit_val);
  free(invalid_verts); free(init_corners); free(tv.valence); free(tv.seams);
  for (int i = 0; i < 6; ++i) free(tv.ctx_syms[i]);
  if (status) live_drop(res);
  return status;
#undef FAIL
}

/* Depth-first attribute traversal over `t` (base or attribute corner table):
 * MeshTraversalSequencer.GenerateSequenceInternal + DepthFirstTraverser.TraverseFromCorner + the observer. */
static void traverse(const view_t *t, uint32_t n_faces, uint32_t *d2c, uint32_t *n_entries, int32_t *v2d) {
  uint8_t *fvis = (uint8_t *)calloc(n_faces ? n_faces : 1, 1);
  uint8_t *vvis = (uint8_t *)calloc(t->n_vertices ? t->n_vertices : 1, 1);
  uint32_t *stk = (uint32_t *)malloc((size_t)(3 * n_faces + 4) * 4);
  uint32_t n = 0;
#define VISIT(v, c) do { vvis[v] = 1; d2c[n] = (c); v2d[v] = (int32_t)n; ++n; } while (0)
  for (uint32_t f = 0; f < n_faces; ++f) {
    uint32_t corner = 3 * f;
    if (fvis[corner / 3]) continue;
    uint32_t sp = 0;
    stk[sp++] = corner;
    uint32_t nv = v_vertex(t, c_next(corner)), pv = v_vertex(t, c_prev(corner));
    if (nv == INV || pv == INV || nv >= t->n_vertices || pv >= t->n_vertices) continue;
    if (!vvis[nv]) VISIT(nv, c_next(corner));
    if (!vvis[pv]) VISIT(pv, c_prev(corner));
    while (sp > 0) {
      corner = stk[sp - 1];
      uint32_t face = corner / 3;
      if (corner == INV || fvis[face]) { --sp; continue; }
      for (;;) {
        fvis[face] = 1;
        uint32_t v = v_vertex(t, corner);
        if (v == INV || v >= t->n_vertices) { sp = 0; break; }
        if (!vvis[v]) {
          int on_b = v_on_boundary(t, v);
          VISIT(v, corner);
          if (!on_b) {
            corner = v_right_corner(t, corner);
            face = corner / 3;
            continue;
          }
        }
        uint32_t rc = v_right_corner(t, corner), lc = v_left_corner(t, corner);
        int rvis = rc == INV || fvis[rc / 3], lvis = lc == INV || fvis[lc / 3];
        if (rvis) {
          if (lvis) { --sp; break; }
          corner = lc; face = lc / 3;
        } else {
          if (lvis) { corner = rc; face = rc / 3; }
          else { stk[sp - 1] = lc; stk[sp++] = rc; break; }
        }
      }
    }
  }
#undef VISIT
  *n_entries = n;
  free(fvis); free(vvis); free(stk);
}

/* per attributes decoder: which tables it traverses / predicts with (CreateAttributesDecoder :640-708,
 * GetAttributeCornerTable :710-731, GetAttributeEncodingData :733-760 with B-15) */
int orc_eb_build_maps(orc_result *res, const uint8_t *dec_ids, int n_dec) {
  eb_state_t *s = g_state_of(res);
  if (!s) return ORC_ERR_CONNECTIVITY;
  const ctab_t *ct = &s->ct;
  int status = ORC_OK;
  res->maps = (orc_mesh_maps *)calloc((size_t)(n_dec ? n_dec : 1), sizeof(orc_mesh_maps));
  res->n_maps = n_dec;
  int pos_used = 0;
  uint8_t *att_used = (uint8_t *)calloc(s->n_attr_data ? s->n_attr_data : 1, 1);
  for (int d = 0; d < n_dec && !status; ++d) {
    int att_data_id = (int8_t)dec_ids[3 * d];
    int dec_type = dec_ids[3 * d + 1];
    int trav_method = dec_ids[3 * d + 2];
    if (trav_method >= 2) { status = ORC_ERR_CONNECTIVITY; break; }
    if (att_data_id >= 0) {
      if ((uint32_t)att_data_id >= s->n_attr_data || att_used[att_data_id]) { status = ORC_ERR_CONNECTIVITY; break; }
      att_used[att_data_id] = 1;
    } else {
      if (pos_used) { status = ORC_ERR_CONNECTIVITY; break; }
      pos_used = 1;
    }
    view_t t;
    uint32_t n_map_verts;
    if (dec_type == 0) { /* vertex attribute: base corner table */
      if (trav_method != 0) { status = ORC_ERR_UNSUPPORTED; break; } /* MaxPredictionDegree: B-16, not restated */
      t.opp = ct->opp; t.c2v = ct->c2v; t.vleft = ct->vcorner; t.n_corners = ct->n_corners; t.n_vertices = ct->n_vertices;
      n_map_verts = ct->n_vertices;
      if (att_data_id >= 0 && s->a_nverts[att_data_id] > n_map_verts) n_map_verts = s->a_nverts[att_data_id];
    } else { /* corner attribute: the attribute's own corner table */
      if (trav_method != 0 || att_data_id < 0) { status = ORC_ERR_CONNECTIVITY; break; }
      t.opp = s->a_opp[att_data_id]; t.c2v = s->a_c2v[att_data_id]; t.vleft = s->a_vleft[att_data_id];
      t.n_corners = ct->n_corners; t.n_vertices = s->a_nverts[att_data_id];
      n_map_verts = t.n_vertices > ct->n_vertices ? t.n_vertices : ct->n_vertices;
    }
    uint32_t *d2c = (uint32_t *)malloc((size_t)(n_map_verts + 4) * 4);
    int32_t *v2d = (int32_t *)calloc((size_t)(n_map_verts ? n_map_verts : 1), 4); /* Resize(n, 0) */
    uint32_t n_entries = 0;
    traverse(&t, ct->n_corners / 3, d2c, &n_entries, v2d);
    uint32_t *opp = (uint32_t *)malloc((size_t)(t.n_corners ? t.n_corners : 1) * 4);
    uint32_t *c2v = (uint32_t *)malloc((size_t)(t.n_corners ? t.n_corners : 1) * 4);
    memcpy(opp, t.opp, (size_t)t.n_corners * 4);
    memcpy(c2v, t.c2v, (size_t)t.n_corners * 4);
    orc_mesh_maps *m = &res->maps[d];
    m->opposite = opp;
    m->corner_to_vertex = c2v;
    m->n_corners = t.n_corners;
    m->data_to_corner = d2c;
    m->n_entries = n_entries;
    m->vertex_to_data = v2d;
    m->n_vertices = n_map_verts;
  }
  free(att_used);
  live_drop(res);
  return status;
}

/* Sequential mesh connectivity (MeshSequentialDecoder.cs:8-118): SURVEY 8f-4, not restated yet. */
/* MeshSequentialDecoder.DecodeConnectivity (D/IO/Mesh/MeshSequentialDecoder.cs:8-83) and DecodeAndDecompressIndices
 * (:85-118), v2.2 container.  Two sites follow the bitstream the C# ports instead of the C# (SURVEY Appendix B):
 * the point count is recorded (the C# never sets it, :23/:122, so its LinearSequencer would cover zero points), and
 * an ODD symbol is a negative index difference (:97 tests `== 0`, the inverse of the encoder it ports; with that test
 * the second index of any ordinary mesh throws).  The range checks keep the C#'s meaning. */
int orc_seq_mesh_connectivity(const uint8_t *buf, uint64_t len, uint64_t *pos, orc_result *res) {
  rd_t r = {buf, len, *pos, 0};
  uint64_t nf = e_varint(&r), np = e_varint(&r);
  uint8_t method = e_u8(&r);
  if (r.err) return r.err;
  if (nf > (1u << 28) || np > 0xFFFFFFFFull) return ORC_ERR_CONNECTIVITY;
  uint32_t *faces = (uint32_t *)malloc((size_t)(nf ? nf : 1) * 12);
  if (!faces) return ORC_ERR_CONNECTIVITY;
  int st = ORC_OK;
  if (method == 0) { /* compressed indices */
    uint32_t *sym = (uint32_t *)malloc((size_t)(nf ? nf : 1) * 12);
    st = orc_decode_symbols(buf, len, &r.pos, (uint32_t)(nf * 3), 1, sym, NULL);
    int32_t last = 0;
    for (uint64_t i = 0; i < nf * 3 && !st; ++i) {
      int32_t diff = (int32_t)(sym[i] >> 1);
      if (sym[i] & 1u) {
        if (diff > last) st = ORC_ERR_CONNECTIVITY; /* :99 would go negative */
        diff = -diff;
      } else if (diff > 0x7FFFFFFF - last) {
        st = ORC_ERR_CONNECTIVITY; /* :107 would overflow */
      }
      last += diff;
      faces[i] = (uint32_t)last;
    }
    free(sym);
  } else if (method == 1) { /* uncompressed indices, width by point count :27-79 */
    for (uint64_t i = 0; i < nf * 3 && !st; ++i) {
      uint64_t v = 0;
      if (np < 256) {
        v = e_u8(&r);
      } else if (np < (1u << 16)) {
        v = e_u8(&r);
        v |= (uint64_t)e_u8(&r) << 8;
      } else if (np < (1u << 21)) {
        v = e_varint(&r);
      } else {
        for (int k = 0; k < 4; ++k) v |= (uint64_t)e_u8(&r) << (8 * k);
      }
      if (r.err) st = r.err;
      faces[i] = (uint32_t)v;
    }
  } else {
    st = ORC_ERR_CONNECTIVITY; /* :81 */
  }
  if (st) {
    free(faces);
    return st;
  }
  res->n_points = (uint32_t)np;
  res->n_faces = (uint32_t)nf;
  res->faces = faces;
  *pos = r.pos;
  return ORC_OK;
}
