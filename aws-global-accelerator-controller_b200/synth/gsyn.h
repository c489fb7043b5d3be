/* gsyn.h — C interface of libgarecon_synth.so (synthetic snapshots for tests and bench.py; see gen.cpp). */
#ifndef GSYN_H
#define GSYN_H
#include <stdint.h>
#include "../../include/garecon.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct gsyn_config {
  uint64_t seed;
  uint32_t n_objects;
  float frac_ingress;
  uint32_t n_zones;
  float frac_r53;
  uint32_t min_hostnames, max_hostnames;
  float frac_wildcard;
  uint32_t svc_ports;
  float frac_hot;
  uint32_t hot_pool;
  float frac_listen_ann;
  float frac_unmanaged;
  float frac_ineligible;
  float p_missing_acc, p_port_drift, p_proto_drift, p_tag_drift, p_missing_listener, p_missing_eg, p_lb_not_active, p_orphan_acc;
  float p_rec_missing, p_alias_drift, p_orphan_rec;
  float p_dup_ports;
  uint32_t intern_keys; /* 1: annotation keys and tag keys are stored once in the slab and referenced by every row (what a
                           packer with a small map[string]ref does); 0: every row carries its own copy */
  /* chunk of a larger cluster (sharded mode): names and random streams use the cluster-wide object index index_base + i and
     zone row zone_base + z; with zones_total > 0 the zone table has zones_total rows (identical on every chunk) and record
     sets only under this chunk's n_zones zones */
  uint32_t index_base, zone_base, zones_total;
  uint32_t layout;    /* 0 = strings row-major by parent (an object's strings are neighbours); 1 = column-major slabs */
  uint32_t emit_mask; /* 0 = all tables; else bit 0 objects, 1 load balancers, 2 accelerators (+ nested), 3 record sets */
  char cluster[64];
} gsyn_config;
typedef struct gsyn_snapshot gsyn_snapshot;
void gsyn_preset(int cfg, uint32_t n, gsyn_config *c);
gsyn_snapshot *gsyn_generate(const gsyn_config *cfg);
const gar_objects *gsyn_objects(const gsyn_snapshot *s);
const gar_actual *gsyn_actual(const gsyn_snapshot *s);
void gsyn_free(gsyn_snapshot *s);
#ifdef __cplusplus
}
#endif
#endif
