// Views of the context's device arrays in the shapes the kernel-body headers take (fc_momentum_body.cuh,
// fc_piso_body.cuh, fc_grad_body.cuh).
#pragma once
#include "fc_internal.cuh"
#include "fc_momentum_body.cuh"

static inline fcm_geom fcm_geom_of(const fc_context *ctx) {
  return fcm_geom{ctx->owner, ctx->neigh, ctx->xc, ctx->yc, ctx->zc, ctx->vol, ctx->arx, ctx->ary, ctx->arz,
                  ctx->xf, ctx->yf, ctx->zf, ctx->facint, ctx->n, ctx->F};
}
static inline fcm_c2f fcm_c2f_of(const fc_context *ctx) {
  return fcm_c2f{ctx->c2f_off, ctx->c2f_face, ctx->c2f_other, ctx->c2f_pos};
}
// first slot / first face / count of inlet, outlet, symmetry, wall, prOutlet; slots follow [cells | npro halo]
static inline fcm_slots fcm_slots_of(const fc_context *ctx) {
  const fc_mesh_desc &m = ctx->m;
  fcm_slots s;
  const int cnt[5] = {m.ninl, m.nout, m.nsym, m.nwal, m.npru};
  const int fst[5] = {m.iInletFacesStart, m.iOutletFacesStart, m.iSymmetryFacesStart, m.iWallFacesStart,
                      m.iPressOutletFacesStart};
  int slot = ctx->n + ctx->npro;
  for (int b = 0; b < 5; ++b) { s.count[b] = cnt[b]; s.face[b] = fst[b]; s.slot[b] = slot; slot += cnt[b]; }
  return s;
}
