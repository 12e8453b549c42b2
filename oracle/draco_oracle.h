/*
 * oracle/draco_oracle.h -- CPU ORACLE. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C restatement of the attribute-decode hot path of B3zaleel/draco-sharp
 * (C#), function by function, with the reference file:line each one follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product library
 * (libdracob200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: see the header of draco_oracle.c.
 */
#ifndef DRACO_ORACLE_H
#define DRACO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* status codes (mirror the reference's exception sites; same numbering as include/dracob200.h) */
enum {
  ORC_OK = 0,
  ORC_ERR_EOF = -1,          /* EndOfStreamException from BinaryReader */
  ORC_ERR_MAGIC = -2,        /* DracoDecoder.cs:47-50 */
  ORC_ERR_UNSUPPORTED = -3,  /* NotImplemented / NotSupported paths (PC kd-tree, out-of-scope predictors, version) */
  ORC_ERR_SCHEME = -4,       /* SymbolDecoding.cs:26 */
  ORC_ERR_BITLEN = -5,       /* SymbolDecoding.cs:56 */
  ORC_ERR_TABLE = -6,        /* RAnsSymbolDecoder.cs:31, RAnsDecoder.cs:80,87 */
  ORC_ERR_RANS_INIT = -7,    /* RAnsDecoder.cs:22,33,39,53 */
  ORC_ERR_PRED = -8,         /* SequentialIntegerAttributeDecoder.cs:26,31 */
  ORC_ERR_WRAP = -9,         /* PredictionSchemeWrapDecodingTransform.cs:73, WrapTransform.cs:91 */
  ORC_ERR_QUANT = -10,       /* AttributeQuantizationTransform.cs:121, OctahedronToolBox.cs:15 */
  ORC_ERR_ATTR = -11,        /* AttributesDecoder.cs:37-39, SequentialNormalAttributeDecoder.cs:14-15 */
  ORC_ERR_TAG = -12,         /* DecoderBuffer.cs:141 (count > 32) */
  ORC_ERR_NUM_SYMBOLS = -13, /* SymbolDecoding.cs:36,59 */
  ORC_ERR_MAPS = -14,        /* mesh attribute without connectivity maps */
  ORC_ERR_CONNECTIVITY = -15 /* Edgebreaker / sequential connectivity failure */
};

typedef struct orc_attr {
  int32_t att_type, data_type, nc, normalized;
  uint32_t unique_id;
  int32_t seq_type;        /* 0 generic 1 integer 2 quantization 3 normals */
  int32_t decoder_id;      /* attributes-decoder this attribute belongs to */
  int32_t pred_method;     /* -2 none, 0 difference, 1 parallelogram ... */
  int32_t transform;       /* -1 none, 0 delta, 1 wrap, 2 oct, 3 oct canonicalized */
  int32_t compressed;      /* u8 */
  int32_t scheme;          /* 0 tagged, 1 raw, -1 n/a */
  int32_t nc_portable;     /* 2 for normals else nc */
  uint32_t n_entries;
  int32_t max_bit_length, precision;
  uint32_t table_symbols;
  uint64_t table_off, payload_off, payload_len; /* absolute byte offsets in the buffer */
  uint64_t bits_off, bits_len;                  /* tagged bit area */
  uint32_t final_state;                         /* rANS state after the last symbol */
  uint64_t leftover;                            /* unread payload bytes after the last symbol */
  int32_t xf_a, xf_b;      /* wrap: min,max ; oct: max_quantized_value, center_value */
  float qmin[4];
  float qrange;
  int32_t qbits;
  uint32_t *symbols;       /* [n_entries*nc_portable] decoded symbols (NULL if uncompressed/generic) */
  int32_t *corr;           /* [n_entries*nc_portable] values after zig-zag (corrections) */
  int32_t *qints;          /* [n_entries*nc_portable] portable (quantized) integers after prediction */
  uint8_t *out;            /* final attribute bytes, tightly packed */
  uint64_t out_bytes;
} orc_attr;

/* connectivity-derived inputs of the parallelogram predictor for ONE attributes-decoder
 * (MeshPredictionSchemeData.cs:5-24) */
typedef struct orc_mesh_maps {
  const uint32_t *opposite;         /* [n_corners]  CornerTable.Opposite (0xFFFFFFFF = invalid) */
  const uint32_t *corner_to_vertex; /* [n_corners]  CornerTable.Vertex */
  uint64_t n_corners;
  const uint32_t *data_to_corner;   /* [n_entries]  EncodedAttributeValueIndexToCornerMap */
  uint64_t n_entries;
  const int32_t *vertex_to_data;    /* [n_vertices] VertexToEncodedAttributeValueIndexMap */
  uint64_t n_vertices;
} orc_mesh_maps;

typedef struct orc_result {
  int32_t status;
  int32_t ver_major, ver_minor, geom_type, method, flags;
  uint32_t n_points, n_faces;
  int32_t n_decoders, n_attrs;
  orc_attr *attrs;
  uint32_t *faces;             /* [3*n_faces] point ids (mesh only) */
  uint64_t attr_section_off;   /* where ATTRIBUTES begins */
  uint64_t end_off;            /* first byte after the last field read */
  /* per attributes-decoder connectivity products (mesh Edgebreaker only; owned) */
  int32_t n_maps;
  orc_mesh_maps *maps;
} orc_result;

/* ---- primitives (each cites the reference in draco_oracle.c) ---- */
int orc_varint(const uint8_t *p, uint64_t len, uint64_t *pos, uint64_t *out);
uint32_t orc_read_bits_lsb(const uint8_t *p, uint64_t len, uint64_t *bitpos, int count, int *err);
int32_t orc_zigzag(uint32_t v);
uint32_t orc_reinterpret_i2u(int32_t v);
uint64_t orc_int_sqrt(uint64_t n);
int orc_rans_precision(int max_bit_length);
/* SYMBOLS(n, nc) at p[*pos..]; advances *pos to the first byte after the symbols field. */
int orc_decode_symbols(const uint8_t *p, uint64_t len, uint64_t *pos, uint32_t num_values, uint32_t nc,
                       uint32_t *out, orc_attr *diag);
void orc_delta_wrap(const int32_t *corr, uint32_t n, int nc, int32_t mn, int32_t mx, int32_t *out);
void orc_delta_oct(const int32_t *corr, uint32_t n, int32_t max_q, int canonical, int32_t *out);
int orc_parallelogram_wrap(const int32_t *corr, uint32_t n, int nc, int32_t mn, int32_t mx,
                           const orc_mesh_maps *m, int32_t *out);
void orc_dequantize(const int32_t *q, uint32_t n, int nc, const float *mn, float range, int bits, float *out);
void orc_oct_to_unit(const int32_t *st, uint32_t n, int bits, float *out);
void orc_narrow(const int32_t *q, uint64_t count, int data_type, uint8_t *out);

/* ---- whole-buffer decode. maps may be NULL (point clouds; Edgebreaker meshes use the
 * oracle's own host connectivity when built in, see orc_eb.c). */
int orc_decode(const uint8_t *buf, uint64_t len, const orc_mesh_maps *maps, int n_maps, orc_result **out);
int orc_decode_ex(const uint8_t *buf, uint64_t len, const orc_mesh_maps *maps, int n_maps, uint64_t attr_section_off,
                  uint32_t n_points, orc_result **out);
void orc_free(orc_result *r);

/* decode + discard, for timing the CPU baseline; returns status, accumulates a checksum of outputs */
int orc_decode_bench(const uint8_t *buf, uint64_t len, uint64_t *points, uint64_t *out_bytes, uint64_t *checksum);
/* FNV-1a 64 of a byte range (used as a checksum-of-checksums in full-size parity tests) */
uint64_t orc_fnv1a(const uint8_t *p, uint64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
