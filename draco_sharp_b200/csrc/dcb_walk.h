// draco_sharp_b200/csrc/dcb_walk.h -- resumable walk over the ATTRIBUTES payload of one .drc buffer.
//
// Host AND device code (one source): the host indexer runs it at dcb_index time; for buffers that
// hold a Tagged attribute the walk stops at the bit area (its length is Sigma bit_length * nc, known
// only after the tags are decoded -- SymbolDecoding.cs:37-49 has no length prefix) and a one-thread-
// per-buffer kernel resumes it on the device once the tag kernel has written bits_total.
//
// It reads header-sized fields only (varints, probability tables, the last <= 4 payload bytes); it
// never decodes symbols.  Wire layout followed (src/Draco/IO/...):
//   PORTABLE          Attributes/SequentialIntegerAttributeDecoder.cs:23-101, SequentialAttributeDecoder.cs:75-86
//   SYMBOLS           Entropy/SymbolDecoding.cs:7-67
//   RANS_TABLE        Entropy/RAnsSymbolDecoder.cs:12-51, Entropy/RAnsDecoder.cs:69-88
//   payload framing   Entropy/RAnsSymbolDecoder.cs:53-59, Entropy/RAnsDecoder.cs:20-54
//   PRED_DATA         Attributes/PredictionSchemes/PredictionSchemeWrapDecodingTransform.cs:69-75,
//                     ...NormalOctahedronCanonicalizedDecodingTransform.cs:80-84, ...NormalOctahedronDecodingTransform.cs:69-76
//   XFORM_PARAMS      Attributes/AttributeQuantizationTransform.cs:110-121, AttributeOctahedronTransform.cs:39-42
//   wire order        Attributes/AttributesDecoder.cs:65-70 (all PORTABLE of a decoder, then all XFORM_PARAMS)
// Deviations where the C# throws or corrupts follow the bitstream, exactly
// as the CPU oracle does, so per-buffer status codes agree.
#pragma once
#include "dcb_internal.h"

struct WalkRd {
  const uint8_t *p;   // arena base
  uint64_t end;       // one past the last byte of the buffer (absolute)
  uint64_t pos;       // absolute
  int err;
};

DCB_HD bool wr_need(WalkRd &r, uint64_t n) {
  if (r.err) return false;
  if (r.pos > r.end || r.end - r.pos < n) {
    r.err = DCB_ERR_EOF;
    return false;
  }
  return true;
}
DCB_HD uint32_t wr_u8(WalkRd &r) { return wr_need(r, 1) ? r.p[r.pos++] : 0u; }
DCB_HD int32_t wr_i8(WalkRd &r) { return (int32_t)(int8_t)wr_u8(r); }
DCB_HD uint32_t wr_u32(WalkRd &r) {
  if (!wr_need(r, 4)) return 0;
  const uint8_t *q = r.p + r.pos;
  r.pos += 4;
  return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
}
DCB_HD float wr_f32(WalkRd &r) {
  union { uint32_t u; float f; } c;
  c.u = wr_u32(r);
  return c.f;
}
// LEB128 unsigned: DecoderBuffer.cs:26-42 (more than 10 bytes is rejected)
DCB_HD uint64_t wr_varint(WalkRd &r) {
  uint64_t result = 0;
  unsigned shift = 0;
  if (r.err) return 0;
  for (int i = 0; i < 10; ++i) {
    if (r.pos >= r.end) { r.err = DCB_ERR_EOF; return 0; }
    const uint32_t b = r.p[r.pos++];
    result |= (uint64_t)(b & 0x7Fu) << shift;
    if (!(b & 0x80u)) return result;
    shift += 7;
  }
  r.err = DCB_ERR_EOF;
  return 0;
}

// RANS_TABLE: validates it and advances past it.  Entropy/RAnsSymbolDecoder.cs:12-51 + RAnsDecoder.cs:69-88
// rec_need[j] (j = 0..3, kA = DCB_REC_KA0 + j): 8-byte units a lane of the bucket-record kernels (dcb_rans_rec.cu) needs
// for this table (RecShape below), or 0xFFFF when the table does not have the shape.
DCB_HD int walk_rans_table(WalkRd &r, int prec_bits, uint32_t &num_symbols, uint32_t &n_active, uint32_t &dense_prefix,
                           uint16_t *narrow_blk, uint16_t *rec_need) {
  num_symbols = 0;
  n_active = 0;
  dense_prefix = 0;  // number of leading symbols that all have non-zero probability
  bool gap = false;
  RecShape shape;
  shape.begin();
  for (int j = 0; j < 4; ++j) rec_need[j] = 0xFFFFu;
  const uint32_t none_blk = (uint32_t)((1ull << prec_bits) >> 7) > 0xFFFFu ? 0xFFFFu : (uint32_t)((1ull << prec_bits) >> 7);
  for (int k = 0; k < 8; ++k) narrow_blk[k] = (uint16_t)none_blk;
  const uint64_t ns = wr_varint(r);
  if (r.err) return r.err;
  if (ns > (r.end - r.pos) * 64u || ns > (1u << 24)) return DCB_ERR_EOF;  // cannot be backed by data
  num_symbols = (uint32_t)ns;
  if (ns == 0) return DCB_OK;
  uint64_t sum = 0;
  for (uint32_t i = 0; i < num_symbols; ++i) {
    const uint32_t pd = wr_u8(r);
    if (r.err) return r.err;
    const uint32_t token = pd & 3u;
    if (token == 3u) {
      const uint32_t off = pd >> 2;
      if (i + off >= num_symbols) return DCB_ERR_TABLE;  // RAnsSymbolDecoder.cs:31
      gap = true;
      i += off;
    } else {
      uint32_t prob = pd >> 2;
      for (uint32_t b = 0; b < token; ++b) prob |= wr_u8(r) << (8 * (b + 1) - 2);
      if (r.err) return r.err;
      if (prob) {
        ++n_active;
        if (!gap) ++dense_prefix;
        for (int k = 1; k <= 7; ++k)
          if (prob < (1u << k) && narrow_blk[k] == (uint16_t)none_blk && (sum >> 7) < none_blk) narrow_blk[k] = (uint16_t)(sum >> 7);
        if (sum + prob <= (1ull << prec_bits)) shape.entry((uint32_t)sum, prob);
      } else {
        gap = true;
      }
      sum += prob;
    }
  }
  if (sum != (1ull << prec_bits)) return DCB_ERR_TABLE;  // RAnsDecoder.cs:80,87
#if !defined(__CUDA_ARCH__) && defined(DCB_DEBUG_SHAPE)
  fprintf(stderr, "[shape] prec=%d ns=%u active=%u end_wide=%u bad_at: s3=%u s4=%u s5=%u s6=%u s7=%u s8=%u\n", prec_bits, num_symbols, n_active,
          shape.end_wide, shape.bad_at[0], shape.bad_at[1], shape.bad_at[2], shape.bad_at[3], shape.bad_at[4], shape.bad_at[5]);
#endif
  if (prec_bits <= 15) {
    const uint32_t prec = 1u << prec_bits;
    for (int j = 0; j < 4; ++j) {
      uint32_t ta, tb, tc;
      if (!shape.cuts(prec, (uint32_t)(DCB_REC_KA0 + j), ta, tb, tc)) continue;
      const uint32_t units = dcb_rec_layout(ta, tb, tc, prec, (uint32_t)(DCB_REC_KA0 + j)).bytes >> 3;
#if !defined(__CUDA_ARCH__) && defined(DCB_DEBUG_SHAPE)
      fprintf(stderr, "[shape]   ka=%d ta=%u tb=%u tc=%u bytes=%u\n", DCB_REC_KA0 + j, ta, tb, tc, units * 8);
#endif
      rec_need[j] = units < 0xFFFFu ? (uint16_t)units : (uint16_t)0xFFFFu;
    }
  }
  return DCB_OK;
}

// varint payload_len + payload; checks the ReadInit conditions.  RAnsSymbolDecoder.cs:53-59, RAnsDecoder.cs:20-54
DCB_HD int walk_rans_payload(WalkRd &r, int prec_bits, uint64_t &off, uint64_t &len) {
  const uint64_t n = wr_varint(r);
  if (r.err) return r.err;
  if (!wr_need(r, n)) return r.err;
  off = r.pos;
  len = n;
  r.pos += n;
  if (n < 1) return DCB_ERR_RANS_INIT;
  const uint8_t *buf = r.p + off;
  const uint32_t tag = (uint32_t)buf[n - 1] >> 6;
  if (n < (uint64_t)tag + 1u) return DCB_ERR_RANS_INIT;
  uint32_t v = 0;
  for (uint32_t i = 0; i <= tag; ++i) v |= (uint32_t)buf[n - 1 - tag + i] << (8 * i);
  v &= (tag == 0) ? 0x3Fu : (tag == 1) ? 0x3FFFu : (tag == 2) ? 0x3FFFFFu : 0x3FFFFFFFu;
  const uint32_t l_base = 4u << prec_bits;
  if (v + l_base >= l_base * 256u) return DCB_ERR_RANS_INIT;  // RAnsDecoder.cs:53
  return DCB_OK;
}

// PRED_DATA of an attribute whose symbols have been located.  Returns status.
// u8 prob_zero | varint size | data of one RAnsBitDecoder block (BitCoders/RAnsBitDecoder.cs:12-24) with the checks of
// AnsDecoder.ReadInit; r.pos ends behind the block.
DCB_HD int walk_rabs_block(WalkRd &r) {
  wr_u8(r);
  const uint64_t nb = wr_varint(r);
  if (r.err) return r.err;
  if (!wr_need(r, nb)) return r.err;
  if (nb < 1) return DCB_ERR_CONNECTIVITY;
  const uint8_t *tail = r.p + r.pos + nb;
  const uint32_t x = (uint32_t)tail[-1] >> 6;
  uint32_t st;
  if (x == 0) st = tail[-1] & 0x3Fu;
  else if (x == 1) { if (nb < 2) return DCB_ERR_CONNECTIVITY; st = ((uint32_t)tail[-2] | ((uint32_t)tail[-1] << 8)) & 0x3FFFu; }
  else if (x == 2) { if (nb < 3) return DCB_ERR_CONNECTIVITY; st = ((uint32_t)tail[-3] | ((uint32_t)tail[-2] << 8) | ((uint32_t)tail[-1] << 16)) & 0x3FFFFFu; }
  else return DCB_ERR_CONNECTIVITY;
  if (st + 4096u >= 4096u * 256u) return DCB_ERR_CONNECTIVITY;
  r.pos += nb;
  return DCB_OK;
}

DCB_HD int walk_pred_data(WalkRd &r, StreamDesc &s, bool has_scheme, bool mesh_scheme) {
  if (!has_scheme) {
    s.recon = RECON_NONE;
    return DCB_OK;
  }
  if (s.transform == XF_WRAP) {
    if (mesh_scheme && s.pred_method == PRED_TEX_COORDS_PORTABLE) {
      // MeshPredictionSchemeTexCoordsPortableDecoder.DecodePredictionData (:68-84): i32 count, then the rABS-coded flags
      // (RAnsBitDecoder.StartDecoding, BitCoders/RAnsBitDecoder.cs:12-24: u8 prob_zero, varint size, data), in front of
      // the transform data.  Located and validated here; tex_chain_kernel decodes them.
      const int32_t no = (int32_t)wr_u32(r);
      if (r.err) return r.err;
      if (no < 0 || (uint64_t)(uint32_t)no > (uint64_t)s.n_entries * s.ncp + 1ull) return DCB_ERR_PRED;  // one flag per entry at most
      s.n_orient = (uint32_t)no;
      s.orient_off = r.pos;
      const int rc = walk_rabs_block(r);
      if (rc) return rc;
    }
    if (mesh_scheme && s.pred_method == PRED_CONSTRAINED_MULTI) {
      // MeshPredictionSchemeConstrainedMultiParallelogramDecoder.DecodeTransformData (:119-141; v2.2: no mode byte): per
      // context a varint flag count and, when it is not zero, one rABS block.  Located and validated here;
      // cmp_flags_kernel decodes them.
      for (int c = 0; c < 4; ++c) {
        const uint64_t nf = wr_varint(r);
        if (r.err) return r.err;
        if (nf > 4ull * (uint64_t)s.n_entries + 4ull) return DCB_ERR_PRED;  // an entry consumes at most four flags
        s.n_crease[c] = (uint32_t)nf;
        s.crease_off[c] = r.pos;
        if (nf > 0) {
          const int rc = walk_rabs_block(r);
          if (rc) return rc;
        }
      }
    }
    s.xf_a = (int32_t)wr_u32(r);
    s.xf_b = (int32_t)wr_u32(r);
    if (r.err) return r.err;
    if (s.xf_a > s.xf_b) return DCB_ERR_WRAP;                    // WrapDecodingTransform.cs:73
    const int64_t diff = (int64_t)s.xf_b - (int64_t)s.xf_a;      // WrapTransform.cs:90-91
    if ((int32_t)diff < 0 || diff >= 2147483647ll) return DCB_ERR_WRAP;
    if (mesh_scheme) {
      if (s.ncp > 4) return DCB_ERR_UNSUPPORTED;  // the mesh chain kernels are built for 1..4 components
      if ((uint64_t)s.n_entries * s.ncp > 0 && !s.has_maps) return DCB_ERR_MAPS;
      if (s.pred_method == PRED_TEX_COORDS_PORTABLE && s.ncp != 2) return DCB_ERR_PRED;  // ...TexCoordsPortableDecoder.cs:51
      s.recon = RECON_PARA_WRAP;
    } else {
      s.recon = RECON_DELTA_WRAP;
    }
  } else {  // octahedron transforms
    s.xf_a = (int32_t)wr_u32(r);
    if (s.transform == XF_OCT_CANON) s.xf_b = (int32_t)wr_u32(r);
    if (r.err) return r.err;
    if (s.xf_a % 2 == 0) return DCB_ERR_QUANT;                    // ...OctahedronTransform.cs:50
    int msb = -1;
    for (uint32_t v = (uint32_t)s.xf_a; v; v >>= 1) ++msb;
    if (msb + 1 < 2 || msb + 1 > 30) return DCB_ERR_QUANT;        // OctahedronToolBox.cs:15
    if (mesh_scheme) {
      if (s.pred_method != PRED_GEOMETRIC_NORMAL) return DCB_ERR_UNSUPPORTED;  // parallelogram on normals: not a Draco combination
      // MeshPredictionSchemeGeometricNormalDecoder.DecodePredictionData (:72-82; v2.2: no mode byte): behind the
      // transform data, one rABS block with a flip bit per entry.  Located and validated here; geo_flips_kernel decodes it.
      s.orient_off = r.pos;
      s.n_orient = s.n_entries;
      const int rc = walk_rabs_block(r);
      if (rc) return rc;
      if (s.n_entries > 0 && !s.has_maps) return DCB_ERR_MAPS;
      s.recon = (s.transform == XF_OCT_CANON) ? RECON_GEO_OCT_CANON : RECON_GEO_OCT;
      return DCB_OK;
    }
    s.recon = (s.transform == XF_OCT_CANON) ? RECON_DELTA_OCT_CANON : RECON_DELTA_OCT;
  }
  return DCB_OK;
}

DCB_HD void walk_scheme_kind(const BufWalk &w, const StreamDesc &s, bool &has_scheme, bool &mesh_scheme, int &err) {
  has_scheme = false;
  mesh_scheme = false;
  err = DCB_OK;
  if (s.pred_method != PRED_NONE) {
    // which scheme objects exist: SequentialIntegerAttributeDecoder.cs:46-51 (wrap only),
    // SequentialNormalAttributeDecoder.cs:19-27 (B-8: both octahedron transforms)
    if (s.seq_type == SEQ_NORMALS)
      has_scheme = (s.transform == XF_OCT || s.transform == XF_OCT_CANON);
    else
      has_scheme = (s.transform == XF_WRAP);
  }
  if (has_scheme && w.geom_type == 1 && w.method == 1) {  // PredictionSchemeDecoderFactory.cs:9-75
    if (s.pred_method == PRED_PARALLELOGRAM || s.pred_method == PRED_TEX_COORDS_PORTABLE ||
        (s.pred_method == PRED_CONSTRAINED_MULTI && s.transform == XF_WRAP) ||
        (s.pred_method == PRED_GEOMETRIC_NORMAL && s.seq_type == SEQ_NORMALS))
      mesh_scheme = true;
    else if (s.pred_method != PRED_DIFFERENCE)
      err = DCB_ERR_UNSUPPORTED;  // multi-parallelogram and deprecated tex coords (pre-2.2 streams)
  }
}

// PORTABLE of one attribute.  Returns 1 when the walk must stop at a Tagged bit area, 0 otherwise
// (errors in r.err / return through *st).
DCB_HD int walk_portable(WalkRd &r, const BufWalk &w, StreamDesc &s, int *st) {
  *st = DCB_OK;
  const uint64_t n = s.n_entries;
  if (s.seq_type == SEQ_GENERIC) {
    const uint64_t bytes = (uint64_t)dcb_dtype_len(s.data_type) * s.nc * n;
    if (!wr_need(r, bytes)) { *st = r.err; return 0; }
    s.scheme = SCHEME_GENERIC;
    s.store = STORE_COPY;
    s.recon = RECON_NONE;
    s.raw_off = r.pos;
    r.pos += bytes;
    s.state = ST_PORTABLE;
    return 0;
  }
  s.pred_method = (int8_t)wr_i8(r);
  if (r.err) { *st = r.err; return 0; }
  if (s.pred_method < -2 || s.pred_method >= PRED_COUNT) { *st = DCB_ERR_PRED; return 0; }
  s.transform = XF_NONE;
  if (s.pred_method != PRED_NONE) {
    s.transform = (int8_t)wr_i8(r);
    if (r.err) { *st = r.err; return 0; }
    if (s.transform < -1 || s.transform >= XF_COUNT) { *st = DCB_ERR_PRED; return 0; }
  }
  bool has_scheme, mesh_scheme;
  int e;
  walk_scheme_kind(w, s, has_scheme, mesh_scheme, e);
  if (e) { *st = e; return 0; }
  const uint64_t nv = n * s.ncp;
  s.compressed = (uint8_t)wr_u8(r);
  if (r.err) { *st = r.err; return 0; }
  s.zigzag = (!has_scheme || s.transform == XF_WRAP) ? 1 : 0;  // B-5
  if (s.compressed > 0) {
    if (nv == 0) {
      s.scheme = SCHEME_EMPTY;
    } else {
      const uint32_t scheme = wr_u8(r);
      if (r.err) { *st = r.err; return 0; }
      if (scheme == SCHEME_TAGGED) {
        s.scheme = SCHEME_TAGGED;
        s.max_bit_length = 5;
        s.prec_bits = (uint8_t)dcb_rans_precision(5);
      } else if (scheme == SCHEME_RAW) {
        s.scheme = SCHEME_RAW;
        const uint32_t mbl = wr_u8(r);
        if (r.err) { *st = r.err; return 0; }
        if (mbl < 1 || mbl > 18) { *st = DCB_ERR_BITLEN; return 0; }
        s.max_bit_length = (uint8_t)mbl;
        s.prec_bits = (uint8_t)dcb_rans_precision((int)mbl);
      } else {
        *st = DCB_ERR_SCHEME;
        return 0;
      }
      s.table_off = r.pos;
      e = walk_rans_table(r, s.prec_bits, s.num_symbols, s.n_active, s.dense_prefix, s.narrow_blk, s.rec_need);
      if (!e && s.num_symbols == 0) e = DCB_ERR_NUM_SYMBOLS;
      if (!e) e = walk_rans_payload(r, s.prec_bits, s.payload_off, s.payload_len);
      if (e) { *st = e; return 0; }
      if (s.scheme == SCHEME_TAGGED) {
        s.bits_off = r.pos;
        s.state = ST_TAGS_PENDING;
        return 1;  // the walk resumes in walk_after_tags
      }
    }
  } else {
    const uint32_t nb = wr_u8(r);
    if (r.err) { *st = r.err; return 0; }
    if (nb > 4) { *st = DCB_ERR_UNSUPPORTED; return 0; }
    if (!wr_need(r, (uint64_t)nb * nv)) { *st = r.err; return 0; }
    s.scheme = SCHEME_UNCOMPRESSED;
    s.raw_num_bytes = (uint8_t)nb;
    s.raw_off = r.pos;
    r.pos += (uint64_t)nb * nv;
  }
  e = walk_pred_data(r, s, has_scheme, mesh_scheme);
  if (e) { *st = e; return 0; }
  s.state = ST_PORTABLE;
  return 0;
}

// XFORM_PARAMS of one attribute
DCB_HD int walk_xform(WalkRd &r, StreamDesc &s) {
  if (s.seq_type == SEQ_QUANTIZATION) {
    for (int c = 0; c < s.nc; ++c) {
      const float f = wr_f32(r);
      if (c < 4) s.q_min[c] = f;
    }
    s.q_range = wr_f32(r);
    s.q_bits = (int32_t)wr_u8(r);
    if (r.err) return r.err;
    if (s.q_bits < 1 || s.q_bits > 30) return DCB_ERR_QUANT;
    s.store = STORE_DEQUANT;
  } else if (s.seq_type == SEQ_NORMALS) {
    s.q_bits = (int32_t)wr_u8(r);
    if (r.err) return r.err;
    if (s.q_bits < 2 || s.q_bits > 30) return DCB_ERR_QUANT;
    s.store = STORE_OCT_UNIT;
  } else if (s.seq_type == SEQ_INTEGER) {
    s.store = STORE_NARROW;
  }
  return DCB_OK;
}

// Continue the walk of one buffer until it is done, fails, or stops at a Tagged bit area.
DCB_HD void walk_continue(const uint8_t *arena, BufWalk &w, StreamDesc *streams) {
  if (w.status != DCB_OK || w.phase == 2) return;
  StreamDesc *ss = streams + w.stream_first;
  WalkRd r;
  r.p = arena;
  r.end = w.end;
  r.pos = w.pos;
  r.err = 0;
  int st = DCB_OK;
  if (w.blocked >= 0) {
    StreamDesc &s = ss[w.blocked];
    if (s.state != ST_TAGS_PENDING) return;
    if (s.status != DCB_OK) {  // tag kernel failed the stream (tag > 32, bit area beyond the buffer)
      w.status = s.status;
      return;
    }
    const uint64_t nbytes = (s.bits_total + 7) >> 3;  // EndBitDecoding: ceil(bits / 8) (B-4)
    r.pos = s.bits_off;
    if (!wr_need(r, nbytes)) { w.status = r.err; return; }
    r.pos += nbytes;
    bool has_scheme, mesh_scheme;
    int e;
    walk_scheme_kind(w, s, has_scheme, mesh_scheme, e);
    if (!e) e = walk_pred_data(r, s, has_scheme, mesh_scheme);
    if (e) { w.status = e; w.pos = r.pos; return; }
    s.state = ST_PORTABLE;
    w.blocked = -1;
    w.cur++;
  }
  while (w.phase != 2) {
    int dec_end = w.dec_start;
    const uint8_t dec = ss[w.dec_start].decoder_id;
    while (dec_end < w.stream_count && ss[dec_end].decoder_id == dec) ++dec_end;
    if (w.phase == 0) {
      while (w.cur < dec_end) {
        const int stop = walk_portable(r, w, ss[w.cur], &st);
        if (st) { w.status = st; w.pos = r.pos; return; }
        if (stop) {
          w.blocked = w.cur;
          w.pos = r.pos;
          return;
        }
        w.cur++;
      }
      w.phase = 1;
      w.cur = w.dec_start;
    }
    while (w.cur < dec_end) {
      st = walk_xform(r, ss[w.cur]);
      if (st) { w.status = st; w.pos = r.pos; return; }
      w.cur++;
    }
    for (int i = w.dec_start; i < dec_end; ++i) {
      // StoreTypedValues dispatch: SequentialIntegerAttributeDecoder.cs:103-140
      if (ss[i].seq_type == SEQ_INTEGER && (ss[i].data_type < DT_INT8 || ss[i].data_type > DT_UINT32)) {
        w.status = DCB_ERR_UNSUPPORTED;
        w.pos = r.pos;
        return;
      }
      ss[i].state = ST_READY;
    }
    w.dec_start = dec_end;
    w.cur = dec_end;
    w.phase = (dec_end >= w.stream_count) ? 2 : 0;
  }
  w.pos = r.pos;
}
