/*
 * fc_oracle_par_piso.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).  Included by fc_oracle_par.c.
 *
 * src-parallel/PISO_multiple_correction.f90, src-parallel/PIMPLE_multiple_correction.f90 and
 * src-parallel/get_rAU_x_UEqnH.f90 with the R ranks advanced in lock step.  Differences from the serial routines that
 * are restated as written:
 *   * processor faces use facefluxmass_piso, apr(i) = can, a(diag) -= can, su -= fmpro (PISO :181-202);
 *   * the reference pressure is pinned on rank iPrefProcess = 0 only (read_input.f90:316, PISO :217-230), with
 *     su(pRefCell) = p(pRefCell) in PISO *and* PIMPLE (the serial PIMPLE uses pp(pRefCell));
 *   * the flux correction (inner faces from the matrix, processor faces from apr) and continuityErrors.h come once,
 *     after the npcor loop (PISO :313-350);
 *   * PIMPLE relaxes as p = urf*pp + (1-urf)*p (PIMPLE :92);
 *   * u, v, w, p are exchanged once, after the corrector loop (PISO :383-386) -- the halo values get_rAU_x_UEqnH reads
 *     in the second corrector are those of the last exchange, not the corrected ones;
 *   * get_rAU_x_UEqnH adds the processor-face terms of ALL THREE components to su (get_rAU_x_UEqnH.f90: the v and w
 *     loops also write `su(ijp)`), and it uses whatever apr holds: the momentum coefficients in the first corrector,
 *     the pressure equation's in the later ones.
 * Parity status: UNPINNED (no stored outputs in the reference); checked by properties in tests/test_oracle_piso.py.
 */

void fco_par_get_rAU_x_UEqnH(fco_rank *R, int nr, fco_uvw *X, const fco_piso_opts *o, double **h) {
  FOR_RANKS {
    const fco_mesh *g = &R[r].g;
    const fco_csr *m = &R[r].m;
    fco_fields *f = &R[r].f;
    fco_uvw *x = &X[r];
    const double *apr = R[r].apr;
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
          A1(s[c], ijp) = A1(s[c], ijp) - A1(h[r], A1(m->icell_jcell, i)) * A1(phio[c], ijn);
          A1(s[c], ijn) = A1(s[c], ijn) - A1(h[r], A1(m->jcell_icell, i)) * A1(phio[c], ijp);
        }
        for (int i = 1; i <= g->npro; ++i) { /* written to su for every component, as in the reference */
          int ijp = A1(g->owner, g->iProcFacesStart + i), ijn = n + i;
          A1(f->su, ijp) = A1(f->su, ijp) - A1(apr, i) * A1(phio[c], ijn);
          A1(f->su, ijp) = A1(f->su, ijp) + A1(apr, i) * A1(phio[c], ijp);
        }
        for (int ijp = 1; ijp <= n; ++ijp) {
          double apotime = A1(f->den, ijp) * A1(g->vol, ijp) / o->timestep;
          double sum = 0.0;
          for (int k = A1(m->ioffset, ijp); k <= A1(m->ioffset, ijp + 1) - 1; ++k) sum = sum + A1(h[r], k);
          double off = sum - A1(h[r], A1(m->diag, ijp));
          A1(s[c], ijp) = A1(s[c], ijp) + (apotime + off) * A1(phio[c], ijp);
        }
      }
      for (int i = 1; i <= g->numInnerFaces; ++i) {
        int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
        A1(s[c], ijp) = A1(s[c], ijp) - A1(h[r], A1(m->icell_jcell, i)) * A1(phi[c], ijn);
        A1(s[c], ijn) = A1(s[c], ijn) - A1(h[r], A1(m->jcell_icell, i)) * A1(phi[c], ijp);
      }
      for (int i = 1; i <= g->npro; ++i) { /* `su(ijp) = su(ijp) - apr(i)*v(ijn)`: su, also for v and w */
        int ijp = A1(g->owner, g->iProcFacesStart + i), ijn = n + i;
        A1(f->su, ijp) = A1(f->su, ijp) - A1(apr, i) * A1(phi[c], ijn);
      }
    }
    for (int i = 1; i <= n; ++i) {
      A1(f->u, i) = A1(x->apu, i) * A1(f->su, i);
      A1(f->v, i) = A1(x->apv, i) * A1(x->sv, i);
      A1(f->w, i) = A1(x->apw, i) * A1(x->sw, i);
    }
  }
}

/* continuityErrors.h of src-parallel: local sums, then global_sum */
static void par_continuity(fco_rank *R, int nr, double *sumLocal, double *global) {
  double *part = (double *)calloc((size_t)nr, sizeof(double)), *part2 = (double *)calloc((size_t)nr, sizeof(double));
  FOR_RANKS {
    const fco_mesh *g = &R[r].g;
    fco_fields *f = &R[r].f;
    for (int i = 0; i < g->numCells; ++i) f->res[i] = 0.0;
    for (int i = 1; i <= g->numInnerFaces; ++i) {
      int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(f->flmass, ijp);
      A1(f->res, ijn) = A1(f->res, ijn) + A1(f->flmass, ijp);
    }
    for (int i = 1; i <= g->npro; ++i) {
      int ijp = A1(g->owner, g->iProcFacesStart + i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(R[r].fmpro, i);
    }
    for (int i = 1; i <= g->ninl; ++i) {
      int ijp = A1(g->owner, g->iInletFacesStart + i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(f->fmi, i);
    }
    for (int i = 1; i <= g->nout; ++i) {
      int ijp = A1(g->owner, g->iOutletFacesStart + i);
      A1(f->res, ijp) = A1(f->res, ijp) - A1(f->fmo, i);
    }
    double sl = 0.0, gl = 0.0;
    for (int i = 1; i <= g->numCells; ++i) sl = sl + fabs(A1(f->res, i));
    for (int i = 1; i <= g->numCells; ++i) gl = gl + A1(f->res, i);
    part[r] = sl; part2[r] = gl;
  }
  *sumLocal = gsum(part, nr);
  *global = gsum(part2, nr);
  free(part); free(part2);
}

/* h[r]: scratch of nnz doubles per rank (module hcoef); pRefCell is a cell of rank 0 */
int fco_par_piso(fco_rank *R, int nr, fco_uvw *X, const fco_piso_opts *o, double **h, fco_piso_report *rep) {
  if (o->pRefCell < 1 || o->pRefCell > R[0].g.numCells) return 3;
  FOR_RANKS memcpy(h[r], R[r].f.a, sizeof(double) * (size_t)R[r].m.nnz); /* h = a */
  fco_calcp_opts co;
  memset(&co, 0, sizeof co);
  co.nigrad = o->nigrad; co.flux_variant = 2; co.const_mflux = o->const_mflux; co.flomas = o->flomas; co.sol = o->sol;
  co.sol.parallel = 1;
  rep->nsolves = 0;
  rep->sumLocalContErr = 0.0; rep->globalContErr = 0.0;
  double **pp = (double **)malloc(sizeof(double *) * (size_t)nr);
  FOR_RANKS pp[r] = R[r].f.pp;
  for (int icorr = 1; icorr <= o->ncorr; ++icorr) {
    fco_par_get_rAU_x_UEqnH(R, nr, X, o, h);
    par_calcp_assemble_v(R, nr, &co, 2); /* grad(U,V,W) incl. exchanges, face loops with facefluxmass_piso, adjustMassFlow */
    { /* reference pressure on rank iPrefProcess = 0 */
      const fco_csr *m = &R[0].m;
      fco_fields *f = &R[0].f;
      for (int k = A1(m->ioffset, o->pRefCell); k <= A1(m->ioffset, o->pRefCell + 1) - 1; ++k) A1(f->a, k) = 0.0;
      A1(f->a, A1(m->diag, o->pRefCell)) = 1.0;
      A1(f->su, o->pRefCell) = A1(f->p, o->pRefCell);
    }
    for (int ipcorr = 1; ipcorr <= o->npcor; ++ipcorr) {
      fco_par_solve(R, nr, 1, pp, &co.sol, &rep->rep[rep->nsolves < 16 ? rep->nsolves : 15], 0); /* iccg(pp,ip) */
      rep->nsolves++;
    }
    FOR_RANKS {
      const fco_mesh *g = &R[r].g;
      const fco_csr *m = &R[r].m;
      fco_fields *f = &R[r].f;
      for (int iface = 1; iface <= g->numInnerFaces; ++iface) {
        int ijp = A1(g->owner, iface), ijn = A1(g->neighbour, iface);
        A1(f->flmass, iface) = A1(f->flmass, iface) + A1(f->a, A1(m->icell_jcell, iface)) * (A1(f->pp, ijn) - A1(f->pp, ijp));
      }
      for (int i = 1; i <= g->npro; ++i) {
        int ijp = A1(g->owner, g->iProcFacesStart + i);
        A1(R[r].fmpro, i) = A1(R[r].fmpro, i) + A1(R[r].apr, i) * (A1(f->pp, g->numCells + i) - A1(f->pp, ijp));
      }
    }
    par_continuity(R, nr, &rep->sumLocalContErr, &rep->globalContErr);
    FOR_RANKS {
      const fco_mesh *g = &R[r].g;
      fco_fields *f = &R[r].f;
      if (o->pimple) {
        for (int inp = 1; inp <= g->numCells; ++inp)
          A1(f->p, inp) = o->urf_p * A1(f->pp, inp) + (1.0 - o->urf_p) * A1(f->p, inp);
      } else {
        for (int i = 0; i < g->numTotal; ++i) f->p[i] = f->pp[i]; /* p = pp */
      }
    }
    for (int istage = 1; istage <= o->nipgrad; ++istage) {
      FOR_RANKS fco_bpres(&R[r].g, R[r].f.p, R[r].f.dPdxi, istage);
      par_grad_field(R, nr, 4, 3, o->nigrad); /* grad(p,dPdxi): exchange(p), Gauss passes, exchange of the gradient */
    }
    FOR_RANKS {
      const fco_mesh *g = &R[r].g;
      fco_fields *f = &R[r].f;
      fco_uvw *x = &X[r];
      for (int inp = 1; inp <= g->numCells; ++inp) {
        A1(f->u, inp) = A1(f->u, inp) - A1(x->apu, inp) * G3(f->dPdxi, 0, inp) * A1(g->vol, inp);
        A1(f->v, inp) = A1(f->v, inp) - A1(x->apv, inp) * G3(f->dPdxi, 1, inp) * A1(g->vol, inp);
        A1(f->w, inp) = A1(f->w, inp) - A1(x->apw, inp) * G3(f->dPdxi, 2, inp) * A1(g->vol, inp);
      }
    }
    par_correct_bc_velocity(R, nr, o->flomas, o->sol.small);
  }
  {
    double **v = (double **)malloc(sizeof(double *) * (size_t)nr);
    for (int c = 0; c < 4; ++c) {
      FOR_RANKS v[r] = c == 0 ? R[r].f.u : c == 1 ? R[r].f.v : c == 2 ? R[r].f.w : R[r].f.p;
      fco_par_exchange(R, nr, v, 1);
    }
    free(v);
  }
  free(pp);
  return 0;
}
