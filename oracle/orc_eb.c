/*
 * oracle/orc_eb.c -- CPU ORACLE (test infrastructure): host connectivity for meshes.
 * Placeholder until the Edgebreaker restatement lands.
 */
#include "draco_oracle.h"
#include <stddef.h>

int orc_eb_decode_connectivity(const uint8_t *buf, uint64_t len, uint64_t *pos, int traversal_type, orc_result *res) {
  (void)buf; (void)len; (void)pos; (void)traversal_type; (void)res;
  return ORC_ERR_UNSUPPORTED;
}
int orc_eb_build_maps(orc_result *res, const uint8_t *dec_ids, int n_dec) {
  (void)res; (void)dec_ids; (void)n_dec;
  return ORC_ERR_UNSUPPORTED;
}
int orc_seq_mesh_connectivity(const uint8_t *buf, uint64_t len, uint64_t *pos, orc_result *res) {
  (void)buf; (void)len; (void)pos; (void)res;
  return ORC_ERR_UNSUPPORTED;
}
