/*
 * fc_oracle_par_uvw.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).  Included by fc_oracle_par.c.
 *
 * src-parallel semantics of the momentum predictor (src-parallel/calcuvw.f90, src-parallel/fieldManipulation.f90
 * calcPressDiv :57-190), R ranks in lock step.  Differences to the serial routine (fc_oracle_uvw.c):
 *   - processor-boundary faces: facefluxuvw with the halo cell as neighbour, apr(i) = can, sp* -= can,
 *     s* += sup.. (calcuvw :225-254); calcPressDiv adds their -fie*S to the owner (:130-146);
 *   - the main diagonal is assembled as a(diag) = sp - a(k1) - a(k2) - ... (running subtraction, :485-493), not
 *     as "sp - (row sum - stale diagonal)";
 *   - Crank-Nicolson: apr = 0.5 apr (:427) and s(ijp) = s(ijp) - apr(i) uo(ijn) + apr(i) uo(ijp) per processor face
 *     (:450-463);
 *   - the solves are src-parallel/bicgstab.f90; at the end exchange(u), (v), (w), (apu) (:657-665).
 * `vis` must hold valid halo values on entry (the reference exchanges it where it is updated).
 * Parity status: UNPINNED (no stored outputs in the reference).
 */

static void par_uvw_faces_rank(fco_rank *Rr, fco_uvw *x, const fco_uvw_opts *o) {
  const fco_mesh *g = &Rr->g;
  const fco_csr *m = &Rr->m;
  fco_fields *f = &Rr->f;
  const int n = g->numCells;
  const int iInletStart = g->numCells + g->npro;
  const int iOutletStart = iInletStart + g->ninl, iSymmetryStart = iOutletStart + g->nout;
  const int iWallStart = iSymmetryStart + g->nsym, iPressOutletStart = iWallStart + g->nwal;
  for (int i = 0; i < n; ++i) { f->su[i] = 0.0; x->sv[i] = 0.0; x->sw[i] = 0.0; x->spu[i] = 0.0; x->spv[i] = 0.0; x->sp[i] = 0.0; }
  /* calcPressDiv: inner faces, processor faces, boundaries */
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    double fie = pres_face_value(g, ijp, ijn, A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->facint, i), f->p, f->dPdxi);
    double dfxe = fie * A1(g->arx, i), dfye = fie * A1(g->ary, i), dfze = fie * A1(g->arz, i);
    A1(f->su, ijp) = A1(f->su, ijp) - dfxe; A1(x->sv, ijp) = A1(x->sv, ijp) - dfye; A1(x->sw, ijp) = A1(x->sw, ijp) - dfze;
    A1(f->su, ijn) = A1(f->su, ijn) + dfxe; A1(x->sv, ijn) = A1(x->sv, ijn) + dfye; A1(x->sw, ijn) = A1(x->sw, ijn) + dfze;
  }
  for (int i = 1; i <= g->npro; ++i) {
    int iface = g->iProcFacesStart + i, ijp = A1(g->owner, iface), ijn = n + i;
    double fie = pres_face_value(g, ijp, ijn, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface), A1(g->fpro, i), f->p, f->dPdxi);
    A1(f->su, ijp) = A1(f->su, ijp) - fie * A1(g->arx, iface);
    A1(x->sv, ijp) = A1(x->sv, ijp) - fie * A1(g->ary, iface);
    A1(x->sw, ijp) = A1(x->sw, ijp) - fie * A1(g->arz, iface);
  }
  {
    const int cnt[5] = {g->ninl, g->nout, g->nsym, g->nwal, g->npru};
    const int fst[5] = {g->iInletFacesStart, g->iOutletFacesStart, g->iSymmetryFacesStart, g->iWallFacesStart,
                        g->iPressOutletFacesStart};
    const int sst[5] = {iInletStart, iOutletStart, iSymmetryStart, iWallStart, iPressOutletStart};
    for (int b = 0; b < 5; ++b)
      for (int i = 1; i <= cnt[b]; ++i) {
        int iface = fst[b] + i, ijp = A1(g->owner, iface), ijb = sst[b] + i;
        A1(f->su, ijp) = A1(f->su, ijp) - A1(f->p, ijb) * A1(g->arx, iface);
        A1(x->sv, ijp) = A1(x->sv, ijp) - A1(f->p, ijb) * A1(g->ary, iface);
        A1(x->sw, ijp) = A1(x->sw, ijp) - A1(f->p, ijb) * A1(g->arz, iface);
      }
  }
  /* volume sources: identical to the serial routine */
  for (int inp = 1; inp <= n; ++inp) {
    if (o->const_mflux) A1(f->su, inp) = A1(f->su, inp) + o->gradPcmf * A1(g->vol, inp);
    if (o->lbuoy) {
      double heat;
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
      A1(f->su, inp) = A1(f->su, inp) + sut; A1(x->sv, inp) = A1(x->sv, inp) + svt; A1(x->sw, inp) = A1(x->sw, inp) + swt;
      A1(x->spu, inp) = A1(x->spu, inp) + apotime * (1 + 0.5 * o->btime);
      A1(x->spv, inp) = A1(x->spv, inp) + apotime * (1 + 0.5 * o->btime);
      A1(x->sp, inp) = A1(x->sp, inp) + apotime * (1 + 0.5 * o->btime);
    }
  }
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    double cap, can, sup, svp, swp;
    fco_facefluxuvw(g, f, x, o, ijp, ijn, A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->arx, i), A1(g->ary, i),
                    A1(g->arz, i), A1(f->flmass, i), A1(g->facint, i), o->gds, &cap, &can, &sup, &svp, &swp);
    A1(f->a, A1(m->icell_jcell, i)) = can;
    A1(f->a, A1(m->jcell_icell, i)) = cap;
    A1(f->su, ijp) = A1(f->su, ijp) + sup; A1(x->sv, ijp) = A1(x->sv, ijp) + svp; A1(x->sw, ijp) = A1(x->sw, ijp) + swp;
    A1(f->su, ijn) = A1(f->su, ijn) - sup; A1(x->sv, ijn) = A1(x->sv, ijn) - svp; A1(x->sw, ijn) = A1(x->sw, ijn) - swp;
  }
  for (int i = 1; i <= g->npro; ++i) { /* :225-254 */
    int iface = g->iProcFacesStart + i, ijp = A1(g->owner, iface), ijn = n + i;
    double cap, can, sup, svp, swp;
    fco_facefluxuvw(g, f, x, o, ijp, ijn, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface), A1(g->arx, iface),
                    A1(g->ary, iface), A1(g->arz, iface), A1(Rr->fmpro, i), A1(g->fpro, i), o->gds, &cap, &can, &sup,
                    &svp, &swp);
    A1(Rr->apr, i) = can;
    A1(x->spu, ijp) = A1(x->spu, ijp) - can; A1(x->spv, ijp) = A1(x->spv, ijp) - can; A1(x->sp, ijp) = A1(x->sp, ijp) - can;
    A1(f->su, ijp) = A1(f->su, ijp) + sup; A1(x->sv, ijp) = A1(x->sv, ijp) + svp; A1(x->sw, ijp) = A1(x->sw, ijp) + swp;
  }
  for (int b = 0; b < 2; ++b) {
    const int cnt = b ? g->nout : g->ninl, fst = b ? g->iOutletFacesStart : g->iInletFacesStart;
    const int sst = b ? iOutletStart : iInletStart;
    const double *fm = b ? f->fmo : f->fmi;
    for (int i = 1; i <= cnt; ++i) {
      int iface = fst + i, ijp = A1(g->owner, iface), ijb = sst + i;
      double cb = facefluxuvw_boundary_can(g, x, ijp, ijb, A1(g->xf, iface), A1(g->yf, iface), A1(g->zf, iface),
                                           A1(g->arx, iface), A1(g->ary, iface), A1(g->arz, iface), A1(fm, i));
      A1(x->spu, ijp) = A1(x->spu, ijp) - cb; A1(x->spv, ijp) = A1(x->spv, ijp) - cb; A1(x->sp, ijp) = A1(x->sp, ijp) - cb;
      A1(f->su, ijp) = A1(f->su, ijp) - cb * A1(f->u, ijb);
      A1(x->sv, ijp) = A1(x->sv, ijp) - cb * A1(f->v, ijb);
      A1(x->sw, ijp) = A1(x->sw, ijp) - cb * A1(f->w, ijb);
    }
  }
  for (int b = 0; b < 2; ++b) {
    const int cnt = b ? g->nwal : g->nsym, fst = b ? g->iWallFacesStart : g->iSymmetryFacesStart;
    const int sst = b ? iWallStart : iSymmetryStart;
    for (int i = 1; i <= cnt; ++i) {
      int iface = fst + i, ijp = A1(g->owner, iface), ijb = sst + i;
      double ax = A1(g->arx, iface), ay = A1(g->ary, iface), az = A1(g->arz, iface);
      double are = sqrt(ax * ax + ay * ay + az * az);
      double nxf = ax / are, nyf = ay / are, nzf = az / are;
      double dn = (A1(g->xf, iface) - A1(g->xc, ijp)) * nxf + (A1(g->yf, iface) - A1(g->yc, ijp)) * nyf +
                  (A1(g->zf, iface) - A1(g->zc, ijp)) * nzf;
      double srd = are / dn;
      double visc = b ? o->viscos : A1(x->vis, ijb);
      double cf = visc * srd;
      double dx = A1(g->xc, ijp) - A1(g->xf, iface), dy = A1(g->yc, ijp) - A1(g->yf, iface),
             dz = A1(g->zc, ijp) - A1(g->zf, iface);
      double dpb = sqrt(dx * dx + dy * dy + dz * dz);
      double vsol = visc * are / dpb;
      double upb = A1(f->u, ijp) - A1(f->u, ijb), vpb = A1(f->v, ijp) - A1(f->v, ijb), wpb = A1(f->w, ijp) - A1(f->w, ijb);
      A1(x->spu, ijp) = A1(x->spu, ijp) + vsol; A1(x->spv, ijp) = A1(x->spv, ijp) + vsol; A1(x->sp, ijp) = A1(x->sp, ijp) + vsol;
      if (!b) {
        double fdne = 2 * cf * (upb * nxf + vpb * nyf + wpb * nzf);
        A1(f->su, ijp) = A1(f->su, ijp) + vsol * A1(f->u, ijp) - fdne * nxf;
        A1(x->sv, ijp) = A1(x->sv, ijp) + vsol * A1(f->v, ijp) - fdne * nyf;
        A1(x->sw, ijp) = A1(x->sw, ijp) + vsol * A1(f->w, ijp) - fdne * nzf;
      } else {
        double vnp = upb * nxf + vpb * nyf + wpb * nzf;
        double utp = upb - vnp * nxf, vtp = vpb - vnp * nyf, wtp = wpb - vnp * nzf;
        A1(f->su, ijp) = A1(f->su, ijp) + vsol * A1(f->u, ijp) - cf * utp;
        A1(x->sv, ijp) = A1(x->sv, ijp) + vsol * A1(f->v, ijp) - cf * vtp;
        A1(x->sw, ijp) = A1(x->sw, ijp) + vsol * A1(f->w, ijp) - cf * wtp;
      }
    }
  }
  if (o->cn) {
    for (int k = 0; k < m->nnz; ++k) f->a[k] = 0.5 * f->a[k];
    for (int i = 0; i < g->npro; ++i) Rr->apr[i] = 0.5 * Rr->apr[i];
  }
}

static void par_uvw_component_rank(fco_rank *Rr, fco_uvw *x, const fco_uvw_opts *o, int comp) {
  const fco_mesh *g = &Rr->g;
  const fco_csr *m = &Rr->m;
  fco_fields *f = &Rr->f;
  const int n = g->numCells;
  double *s = comp == 0 ? f->su : comp == 1 ? x->sv : x->sw;
  double *spc = comp == 0 ? x->spu : comp == 1 ? x->spv : x->sp;
  double *phi = comp == 0 ? f->u : comp == 1 ? f->v : f->w;
  const double *phio = comp == 0 ? x->uo : comp == 1 ? x->vo : x->wo;
  double *ap = comp == 0 ? x->apu : comp == 1 ? x->apv : x->apw;
  if (o->cn) {
    for (int i = 1; i <= g->numInnerFaces; ++i) {
      int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
      A1(s, ijp) = A1(s, ijp) - A1(f->a, A1(m->icell_jcell, i)) * A1(phio, ijn);
      A1(s, ijn) = A1(s, ijn) - A1(f->a, A1(m->jcell_icell, i)) * A1(phio, ijp);
    }
    for (int i = 1; i <= g->npro; ++i) { /* :450-463 */
      int ijp = A1(g->owner, g->iProcFacesStart + i), ijn = n + i;
      A1(s, ijp) = A1(s, ijp) - A1(Rr->apr, i) * A1(phio, ijn);
      A1(s, ijp) = A1(s, ijp) + A1(Rr->apr, i) * A1(phio, ijp);
    }
    for (int ijp = 1; ijp <= n; ++ijp) {
      double apotime = A1(f->den, ijp) * A1(g->vol, ijp) / o->timestep;
      double sum = 0.0;
      for (int k = A1(m->ioffset, ijp); k <= A1(m->ioffset, ijp + 1) - 1; ++k) sum = sum + A1(f->a, k);
      double off = sum - A1(f->a, A1(m->diag, ijp));
      A1(s, ijp) = A1(s, ijp) + (apotime + off) * A1(phio, ijp);
      A1(spc, ijp) = A1(spc, ijp) + apotime;
    }
  }
  const double urfrs = 1.0 / o->urf[comp], urfms = 1.0 - o->urf[comp];
  if (comp > 0)
    for (int inp = 1; inp <= n; ++inp) { A1(f->a, A1(m->diag, inp)) = 0.0; A1(f->su, inp) = 0.0; }
  for (int inp = 1; inp <= n; ++inp) { /* :485-500 */
    double d = A1(spc, inp);
    for (int k = A1(m->ioffset, inp); k <= A1(m->ioffset, inp + 1) - 1; ++k) {
      if (k == A1(m->diag, inp)) continue;
      d = d - A1(f->a, k);
    }
    d = d * urfrs;
    A1(f->a, A1(m->diag, inp)) = d;
    A1(f->su, inp) = A1(s, inp) + urfms * d * A1(phi, inp);
    A1(ap, inp) = 1.0 / (d + o->sol.small);
  }
}

/* gradient of field `which` (0 u, 1 v, 2 w, 4 p) into its gradient array, src-parallel/gradients.f90 */
static void par_grad_named(fco_rank *R, int nr, int which, int nigrad) {
  double **phi = (double **)malloc(sizeof(double *) * (size_t)nr), **gr = (double **)malloc(sizeof(double *) * (size_t)nr);
  FOR_RANKS {
    fco_fields *f = &R[r].f;
    phi[r] = which == 0 ? f->u : which == 1 ? f->v : which == 2 ? f->w : f->p;
    gr[r] = which == 0 ? f->dUdxi : which == 1 ? f->dVdxi : which == 2 ? f->dWdxi : f->dPdxi;
  }
  fco_par_grad_gauss(R, nr, phi, nigrad, gr);
  free(phi); free(gr);
}

void fco_par_calcuvw_assemble(fco_rank *R, int nr, fco_uvw *X, const fco_uvw_opts *o) {
  par_grad_named(R, nr, 0, o->nigrad);
  par_grad_named(R, nr, 1, o->nigrad);
  par_grad_named(R, nr, 2, o->nigrad);
  for (int istage = 1; istage <= o->nipgrad; ++istage) {
    FOR_RANKS fco_bpres(&R[r].g, R[r].f.p, R[r].f.dPdxi, istage);
    par_grad_named(R, nr, 4, o->nigrad);
  }
  FOR_RANKS par_uvw_faces_rank(&R[r], &X[r], o);
}

int fco_par_calcuvw_component(fco_rank *R, int nr, fco_uvw *X, const fco_uvw_opts *o, int comp, fco_report *rep) {
  FOR_RANKS par_uvw_component_rank(&R[r], &X[r], o, comp);
  double **fi = (double **)malloc(sizeof(double *) * (size_t)nr);
  FOR_RANKS fi[r] = comp == 0 ? R[r].f.u : comp == 1 ? R[r].f.v : R[r].f.w;
  fco_solver_opts so = o->sol;
  so.sor = o->sor[comp];
  so.nsw = o->nsw[comp];
  so.parallel = 1;
  int rc = fco_par_solve(R, nr, 2, fi, &so, rep, 0);
  free(fi);
  return rc;
}

int fco_par_calcuvw(fco_rank *R, int nr, fco_uvw *X, const fco_uvw_opts *o, fco_uvw_report *rep) {
  FOR_RANKS if (R[r].g.noc > 0) return 2;
  fco_par_calcuvw_assemble(R, nr, X, o);
  for (int comp = 0; comp < 3; ++comp) {
    int rc = fco_par_calcuvw_component(R, nr, X, o, comp, &rep->rep[comp]);
    if (rc) return rc;
  }
  double **v = (double **)malloc(sizeof(double *) * (size_t)nr);
  for (int c = 0; c < 4; ++c) { /* :657-665 */
    FOR_RANKS v[r] = c == 0 ? R[r].f.u : c == 1 ? R[r].f.v : c == 2 ? R[r].f.w : X[r].apu;
    fco_par_exchange(R, nr, v, 1);
  }
  free(v);
  return 0;
}
