/*
 * fc_oracle_piso.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).  Included by fc_oracle.c.
 *
 * CPU restatement of the PISO / PIMPLE pressure equation:
 *   PISO_multiple_correction    src/PISO_multiple_correction.f90:2-312
 *   PIMPLE_multiple_correction  src/PIMPLE_multiple_correction.f90 (differences: su(pRefCell) = pp(pRefCell) :180,
 *                               one flux correction / continuity report after the npcor loop :260-279,
 *                               p = p + urf(ip) (pp - p) on the cells :282-284)
 *   get_rAU_x_UEqnH             src/get_rAU_x_UEqnH.f90:2-208
 * Serial `src` semantics, no O-C cuts.  It re-uses facefluxmass_piso (variant 2), adjustMassFlow, bpres, the
 * Gauss gradient, iccg and correctBoundaryConditionsVelocity of the pressure-correction oracle.
 *
 * Parity status: UNPINNED (the reference stores no outputs of these routines); checked by properties in
 * tests/test_oracle_piso.py.
 *
 * H(u) of get_rAU_x_UEqnH keeps only the unsteady, buoyancy and neighbour terms (get_rAU_x_UEqnH.f90:24-200): the
 * wall-shear / inlet / deferred-correction sources of calcuvw are not carried over.  Restated as written.
 * Quirks kept: only the ROW of pRefCell is cleared (the column entries stay, so iccg works on a matrix that
 * is no longer symmetric, PISO :188-192); the final flux correction reads a(icell_jcell) from that matrix, so
 * faces owned by pRefCell get no correction (:262); pp is not reset between correctors (:199-200);
 * the velocity correction multiplies in the order apu * dPdx * vol (:299-301), not calcp's dPdx * vol * apu.
 */

/* get_rAU_x_UEqnH: u,v,w <- ap* . H(u,v,w) without the pressure gradient; h = momentum matrix (off-diagonals) */
void fco_get_rAU_x_UEqnH(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_piso_opts *o,
                         const double *h) {
  const int n = g->numCells;
  for (int i = 0; i < n; ++i) { f->su[i] = 0.0; x->sv[i] = 0.0; x->sw[i] = 0.0; }
  for (int inp = 1; inp <= n; ++inp) {
    if (o->lbuoy) {
      double heat = 0.0;
      if (o->boussinesq) heat = o->beta * o->densit * (A1(x->t, inp) - o->tref) * A1(g->vol, inp);
      else heat = (o->densit - A1(f->den, inp)) * A1(g->vol, inp);
      A1(f->su, inp) = A1(f->su, inp) - o->gravx * heat;
      A1(x->sv, inp) = A1(x->sv, inp) - o->gravy * heat;
      A1(x->sw, inp) = A1(x->sw, inp) - o->gravz * heat;
    }
    if (o->bdf) {
      double apotime = A1(f->den, inp) * A1(g->vol, inp) / o->timestep;
      double sut = apotime * ((1 + o->btime) * A1(x->uo, inp));
      double svt = apotime * ((1 + o->btime) * A1(x->vo, inp));
      double swt = apotime * ((1 + o->btime) * A1(x->wo, inp));
      if (o->btime > (double)0.99f) {
        sut = sut - apotime * (0.5 * o->btime * A1(x->uoo, inp));
        svt = svt - apotime * (0.5 * o->btime * A1(x->voo, inp));
        swt = swt - apotime * (0.5 * o->btime * A1(x->woo, inp));
      }
      A1(f->su, inp) = A1(f->su, inp) + sut;
      A1(x->sv, inp) = A1(x->sv, inp) + svt;
      A1(x->sw, inp) = A1(x->sw, inp) + swt;
    }
  }
  double *s[3] = {f->su, x->sv, x->sw};
  const double *phi[3] = {f->u, f->v, f->w}, *phio[3] = {x->uo, x->vo, x->wo};
  for (int c = 0; c < 3; ++c) {
    if (o->cn) {
      for (int i = 1; i <= g->numInnerFaces; ++i) {
        int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
        A1(s[c], ijp) = A1(s[c], ijp) - A1(h, A1(m->icell_jcell, i)) * A1(phio[c], ijn);
        A1(s[c], ijn) = A1(s[c], ijn) - A1(h, A1(m->jcell_icell, i)) * A1(phio[c], ijp);
      }
      for (int ijp = 1; ijp <= n; ++ijp) {
        double apotime = A1(f->den, ijp) * A1(g->vol, ijp) / o->timestep;
        double sum = 0.0;
        for (int k = A1(m->ioffset, ijp); k <= A1(m->ioffset, ijp + 1) - 1; ++k) sum = sum + A1(h, k);
        double off = sum - A1(h, A1(m->diag, ijp));
        A1(s[c], ijp) = A1(s[c], ijp) + (apotime + off) * A1(phio[c], ijp);
      }
    }
    for (int i = 1; i <= g->numInnerFaces; ++i) {
      int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
      A1(s[c], ijp) = A1(s[c], ijp) - A1(h, A1(m->icell_jcell, i)) * A1(phi[c], ijn);
      A1(s[c], ijn) = A1(s[c], ijn) - A1(h, A1(m->jcell_icell, i)) * A1(phi[c], ijp);
    }
  }
  for (int i = 1; i <= n; ++i) {
    A1(f->u, i) = A1(x->apu, i) * A1(f->su, i);
    A1(f->v, i) = A1(x->apv, i) * A1(x->sv, i);
    A1(f->w, i) = A1(x->apw, i) * A1(x->sw, i);
  }
}

int fco_piso(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_piso_opts *o, double *h,
             fco_piso_report *rep) {
  const int n = g->numCells;
  if (g->noc > 0 || g->npro > 0) return 2;
  if (o->pRefCell < 1 || o->pRefCell > n) return 3;
  memcpy(h, f->a, sizeof(double) * (size_t)m->nnz); /* h = a */
  fco_calcp_opts co;
  memset(&co, 0, sizeof co);
  co.nigrad = o->nigrad; co.flux_variant = 2; co.const_mflux = o->const_mflux; co.flomas = o->flomas; co.sol = o->sol;
  rep->nsolves = 0;
  rep->sumLocalContErr = 0.0; rep->globalContErr = 0.0;
  for (int icorr = 1; icorr <= o->ncorr; ++icorr) {
    fco_get_rAU_x_UEqnH(g, m, f, x, o, h);
    /* grad(U,V,W), a = 0, su = 0, facefluxmass_piso face loop, adjustMassFlow (:104-181) */
    fco_calcp_assemble(g, m, f, &co);
    /* reference pressure (:188-192) */
    for (int k = A1(m->ioffset, o->pRefCell); k <= A1(m->ioffset, o->pRefCell + 1) - 1; ++k) A1(f->a, k) = 0.0;
    A1(f->a, A1(m->diag, o->pRefCell)) = 1.0;
    A1(f->su, o->pRefCell) = o->pimple ? A1(f->pp, o->pRefCell) : A1(f->p, o->pRefCell);
    for (int ipcorr = 1; ipcorr <= o->npcor; ++ipcorr) {
      fco_report *r = &rep->rep[rep->nsolves < 16 ? rep->nsolves : 15];
      fco_iccg(m, f->a, f->su, f->pp, f->res, 0, &o->sol, r, 0);
      rep->nsolves++;
      if (!o->pimple) {
        if (ipcorr == o->npcor)
          for (int iface = 1; iface <= g->numInnerFaces; ++iface) {
            int ijp = A1(g->owner, iface), ijn = A1(g->neighbour, iface);
            A1(f->flmass, iface) = A1(f->flmass, iface) + A1(f->a, A1(m->icell_jcell, iface)) * (A1(f->pp, ijn) - A1(f->pp, ijp));
          }
        continuity_errors(g, f, &rep->sumLocalContErr, &rep->globalContErr);
      }
    }
    if (o->pimple) {
      for (int iface = 1; iface <= g->numInnerFaces; ++iface) {
        int ijp = A1(g->owner, iface), ijn = A1(g->neighbour, iface);
        A1(f->flmass, iface) = A1(f->flmass, iface) + A1(f->a, A1(m->icell_jcell, iface)) * (A1(f->pp, ijn) - A1(f->pp, ijp));
      }
      continuity_errors(g, f, &rep->sumLocalContErr, &rep->globalContErr);
      for (int inp = 1; inp <= n; ++inp) A1(f->p, inp) = A1(f->p, inp) + o->urf_p * (A1(f->pp, inp) - A1(f->p, inp));
    } else {
      for (int i = 0; i < g->numTotal; ++i) f->p[i] = f->pp[i]; /* p = pp */
    }
    for (int istage = 1; istage <= o->nipgrad; ++istage) {
      fco_bpres(g, f->p, f->dPdxi, istage);
      fco_grad(g, m, f->p, o->nigrad, f->dPdxi);
    }
    for (int inp = 1; inp <= n; ++inp) {
      A1(f->u, inp) = A1(f->u, inp) - A1(x->apu, inp) * G3(f->dPdxi, 0, inp) * A1(g->vol, inp);
      A1(f->v, inp) = A1(f->v, inp) - A1(x->apv, inp) * G3(f->dPdxi, 1, inp) * A1(g->vol, inp);
      A1(f->w, inp) = A1(f->w, inp) - A1(x->apw, inp) * G3(f->dPdxi, 2, inp) * A1(g->vol, inp);
    }
    correctBoundaryConditionsVelocity(g, f, o->flomas, o->sol.small);
  }
  return 0;
}
