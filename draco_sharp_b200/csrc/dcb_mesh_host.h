// draco_sharp_b200/csrc/dcb_mesh_host.h -- host Edgebreaker connectivity helper (see dcb_mesh_host.cu)
#pragma once
#include <stdint.h>

#include <vector>

struct DcbHostMaps {
  std::vector<uint32_t> opposite, corner_to_vertex, data_to_corner;
  std::vector<int32_t> vertex_to_data;
};

// Decodes the Edgebreaker connectivity of one mesh buffer on the CPU.  Outputs where ATTRIBUTES starts, the
// point count, one set of maps per attributes decoder and (optionally) the faces as 3 point ids each.
int dcb_host_edgebreaker(const uint8_t *buf, uint64_t len, uint64_t conn_off, uint64_t *attr_section_off,
                         uint32_t *n_points, std::vector<DcbHostMaps> *maps, std::vector<uint32_t> *faces);

// Decodes the connectivity of a SEQUENTIAL mesh buffer (MeshSequentialDecoder.cs:8-118) on the CPU: where ATTRIBUTES
// starts, the point count, the faces as 3 point ids each.  Its attributes need no maps.
int dcb_host_sequential(const uint8_t *buf, uint64_t len, uint64_t conn_off, uint64_t *attr_section_off, uint32_t *n_points,
                        std::vector<uint32_t> *faces);
