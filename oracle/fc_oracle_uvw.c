/*
 * fc_oracle_uvw.c -- TEST INFRASTRUCTURE ONLY (see fc_oracle.h).  Included by fc_oracle.c.
 *
 * CPU restatement of the momentum predictor, the caller immediately before the pressure-correction
 * path:
 *   calcuvw                  src/calcuvw.f90:3-557
 *   facefluxuvw / _boundary  src/faceflux_velocity.f90:37-196, 385-549
 *   sngrad ('skewness')      src/gradients.f90:547-668
 *   face_value + schemes     src/interpolation.f90:19-403
 *   calcPressDiv             src/fieldManipulation.f90:57-165, presFaceDivInner :395-445
 * Serial `src` semantics, laminar (lturb = .false.: no calcstress, viss = viscos at walls), no O-C cuts.
 *
 * Parity status: UNPINNED -- the reference stores no outputs of calcuvw.  Checked analytically
 * (tests/test_oracle_momentum.py) and read line by line against the Fortran.
 *
 * Quirks kept on purpose:
 *  - presFaceDivInner indexes its (3,numCells) gradient as df(ijp,1), df(ijp,2), df(ijp,3)
 *    (fieldManipulation.f90:433-435): column-major address arithmetic turns that into the flat elements
 *    ijp, ijp+3, ijp+6 of the array, i.e. components of OTHER cells.  gfortran -O2 has no bounds check,
 *    so this is what the reference computes; it only matters on skewed meshes (the term multiplies
 *    x_f - x_i).  Needs numCells >= 3 for the largest index to stay inside the array.
 *  - face_value receives fxp as `lambda` for the p->n direction (faceflux_velocity.f90:162) and uses it
 *    as the NEIGHBOUR's weight (interpolation.f90:81-85).
 *  - the U diagonal is "row sum minus the stale a(diag)" left by the previous solve
 *    (calcuvw.f90:423-424): the stale value takes part in the floating-point sum.
 *  - the inlet/outlet sup/svp/swp of facefluxuvw_boundary are computed and dropped (:231-241).
 */

#define FCO_MAX2(a, b) (((a) > (b)) ? (a) : (b))
#define FCO_MIN2(a, b) (((a) < (b)) ? (a) : (b))

/* sngrad_scalar_field, approach 'skewness', nrelax = 0 (gradients.f90:595-668) */
static void sngrad_skew(const fco_mesh *g, int ijp, int ijn, double arx, double ary, double arz, double lambda,
                        const double *fi, const double *dFidxi, double *dfixi, double *dfiyi, double *dfizi,
                        double *dfixii, double *dfiyii, double *dfizii) {
  double fxn = lambda, fxp = 1.0 - lambda;
  double xpn = A1(g->xc, ijn) - A1(g->xc, ijp);
  double ypn = A1(g->yc, ijn) - A1(g->yc, ijp);
  double zpn = A1(g->zc, ijn) - A1(g->zc, ijp);
  double costn = 1.0;
  double vole = xpn * arx + ypn * ary + zpn * arz;
  *dfixi = G3(dFidxi, 0, ijp) * fxp + G3(dFidxi, 0, ijn) * fxn;
  *dfiyi = G3(dFidxi, 1, ijp) * fxp + G3(dFidxi, 1, ijn) * fxn;
  *dfizi = G3(dFidxi, 2, ijp) * fxp + G3(dFidxi, 2, ijn) * fxn;
  double d1x = costn, d1y = costn, d1z = costn;
  double d2x = xpn * costn, d2y = ypn * costn, d2z = zpn * costn;
  double rem = A1(fi, ijn) - A1(fi, ijp) - *dfixi * d2x - *dfiyi * d2y - *dfizi * d2z;
  *dfixii = *dfixi * d1x + arx / vole * rem;
  *dfiyii = *dfiyi * d1y + ary / vole * rem;
  *dfizii = *dfizi * d1z + arz / vole * rem;
}

/* face_value and its schemes (interpolation.f90:19-403).  scheme: 0 cds, 1 cdsc, 2 central-f,
 * 3 linear-f (2nd upwind), 4 muscl-f (also the fall-through default), 5 flux limiter with
 * limiter: 0 smart, 1 avl, 2 muscl, 3 umist, 4 koren, 5 charm, 6 ospre, 7 luds (psi = 1). */
static double face_value(const fco_mesh *g, int scheme, int limiter, int ijp, int ijn, double xf, double yf,
                         double zf, double lambda, const double *u, const double *dUdxi) {
  if (scheme == 0) {
    double fxn = lambda, fxp = 1.0 - lambda;
    return A1(u, ijp) * fxp + A1(u, ijn) * fxn;
  }
  if (scheme == 1) {
    double fxn = lambda, fxp = 1.0 - lambda;
    double xi = A1(g->xc, ijp) * fxp + A1(g->xc, ijn) * fxn;
    double yi = A1(g->yc, ijp) * fxp + A1(g->yc, ijn) * fxn;
    double zi = A1(g->zc, ijp) * fxp + A1(g->zc, ijn) * fxn;
    double dfixi = G3(dUdxi, 0, ijp) * fxp + G3(dUdxi, 0, ijn) * fxn;
    double dfiyi = G3(dUdxi, 1, ijp) * fxp + G3(dUdxi, 1, ijn) * fxn;
    double dfizi = G3(dUdxi, 2, ijp) * fxp + G3(dUdxi, 2, ijn) * fxn;
    return A1(u, ijp) * fxp + A1(u, ijn) * fxn + (dfixi * (xf - xi) + dfiyi * (yf - yi) + dfizi * (zf - zi));
  }
  if (scheme == 2) return face_value_central(g, ijp, ijn, xf, yf, zf, u, dUdxi);
  if (scheme == 3) {
    double gradfidr = G3(dUdxi, 0, ijp) * (xf - A1(g->xc, ijp)) + G3(dUdxi, 1, ijp) * (yf - A1(g->yc, ijp)) +
                      G3(dUdxi, 2, ijp) * (zf - A1(g->zc, ijp));
    return A1(u, ijp) + gradfidr;
  }
  if (scheme == 5) {
    double fxp = 1.0 - lambda;
    double xpn = A1(g->xc, ijn) - A1(g->xc, ijp);
    double ypn = A1(g->yc, ijn) - A1(g->yc, ijp);
    double zpn = A1(g->zc, ijn) - A1(g->zc, ijp);
    double r = (2 * G3(dUdxi, 0, ijp) * xpn + 2 * G3(dUdxi, 1, ijp) * ypn + 2 * G3(dUdxi, 2, ijp) * zpn) /
                   (A1(u, ijn) - A1(u, ijp)) - 1.0;
    double psi;
    switch (limiter) {
      case 0: psi = FCO_MAX2(0.0, FCO_MIN2(FCO_MIN2(2.0 * r, 0.75 * r + 0.25), 4.0)); break;
      case 1: psi = FCO_MAX2(0.0, FCO_MIN2(FCO_MIN2(1.5 * r, 0.75 * r + 0.25), 2.5)); break;
      case 2: psi = FCO_MAX2(0.0, FCO_MIN2(FCO_MIN2(2.0 * r, 0.5 * r + 0.5), 2.0)); break;
      case 3: psi = FCO_MAX2(0.0, FCO_MIN2(FCO_MIN2(FCO_MIN2(2.0 * r, 0.75 * r + 0.25), 0.25 * r + 0.75), 2.0)); break;
      case 4: psi = FCO_MAX2(0.0, FCO_MIN2(FCO_MIN2(2.0 * r, 2.0 / 3.0 * r + 1.0 / 3.0), 2.0)); break;
      case 5: psi = (r + fabs(r)) * (3 * r + 1.0) / (2 * ((r + 1.0) * (r + 1.0))); break;
      case 6: psi = 1.5 * r * (r + 1.0) / (r * r + r + 1.0); break;
      default: psi = 1.0; break;
    }
    return A1(u, ijp) + fxp * psi * (A1(u, ijn) - A1(u, ijp));
  }
  { /* muscl-f (interpolation.f90:258-318) */
    double theta = 0.125;
    double up = G3(dUdxi, 0, ijp) * (xf - A1(g->xc, ijp)) + G3(dUdxi, 1, ijp) * (yf - A1(g->yc, ijp)) +
                G3(dUdxi, 2, ijp) * (zf - A1(g->zc, ijp));
    double ce = G3(dUdxi, 0, ijp) * (xf - A1(g->xc, ijp)) + G3(dUdxi, 1, ijp) * (yf - A1(g->yc, ijp)) +
                G3(dUdxi, 2, ijp) * (zf - A1(g->zc, ijp)) + G3(dUdxi, 0, ijn) * (xf - A1(g->xc, ijn)) +
                G3(dUdxi, 1, ijn) * (yf - A1(g->yc, ijn)) + G3(dUdxi, 2, ijn) * (zf - A1(g->zc, ijn));
    double fv_up = (A1(u, ijp) + up);
    double fv_ce = 0.5 * (A1(u, ijp) + A1(u, ijn) + ce);
    return theta * fv_ce + (1.0 - theta) * fv_up;
  }
}

/* facefluxuvw, inner faces (faceflux_velocity.f90:37-196) */
void fco_facefluxuvw(const fco_mesh *g, const fco_fields *f, const fco_uvw *x, const fco_uvw_opts *o, int ijp,
                     int ijn, double xf, double yf, double zf, double arx, double ary, double arz, double flomass,
                     double lambda, double gam, double *cap, double *can, double *sup, double *svp, double *swp) {
  double fxn = lambda, fxp = 1.0 - lambda;
  double xpn = A1(g->xc, ijn) - A1(g->xc, ijp);
  double ypn = A1(g->yc, ijn) - A1(g->yc, ijp);
  double zpn = A1(g->zc, ijn) - A1(g->zc, ijp);
  double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  double are = sqrt(arx * arx + ary * ary + arz * arz);
  double game = A1(x->vis, ijp) * fxp + A1(x->vis, ijn) * fxn;
  double de = game * are / dpn;
  *can = -de + FCO_MIN2(flomass, 0.0);
  *cap = -de - FCO_MAX2(flomass, 0.0);
  double duxi, duyi, duzi, dvxi, dvyi, dvzi, dwxi, dwyi, dwzi;
  double duxii, duyii, duzii, dvxii, dvyii, dvzii, dwxii, dwyii, dwzii;
  sngrad_skew(g, ijp, ijn, arx, ary, arz, lambda, f->u, f->dUdxi, &duxi, &duyi, &duzi, &duxii, &duyii, &duzii);
  sngrad_skew(g, ijp, ijn, arx, ary, arz, lambda, f->v, f->dVdxi, &dvxi, &dvyi, &dvzi, &dvxii, &dvyii, &dvzii);
  sngrad_skew(g, ijp, ijn, arx, ary, arz, lambda, f->w, f->dWdxi, &dwxi, &dwyi, &dwzi, &dwxii, &dwyii, &dwzii);
  double fdue = game * ((duxii + duxii) * arx + (duyii + dvxii) * ary + (duzii + dwxii) * arz);
  double fdve = game * ((duyii + dvxii) * arx + (dvyii + dvyii) * ary + (dvzii + dwyii) * arz);
  double fdwe = game * ((duzii + dwxii) * arx + (dwyii + dvzii) * ary + (dwzii + dwzii) * arz);
  double fdui = game * are / dpn * (duxi * xpn + duyi * ypn + duzi * zpn);
  double fdvi = game * are / dpn * (dvxi * xpn + dvyi * ypn + dvzi * zpn);
  double fdwi = game * are / dpn * (dwxi * xpn + dwyi * ypn + dwzi * zpn);
  double fuuds = FCO_MAX2(flomass, 0.0) * A1(f->u, ijp) + FCO_MIN2(flomass, 0.0) * A1(f->u, ijn);
  double fvuds = FCO_MAX2(flomass, 0.0) * A1(f->v, ijp) + FCO_MIN2(flomass, 0.0) * A1(f->v, ijn);
  double fwuds = FCO_MAX2(flomass, 0.0) * A1(f->w, ijp) + FCO_MIN2(flomass, 0.0) * A1(f->w, ijn);
  double ue, ve, we;
  if (flomass >= 0.0) {
    ue = face_value(g, o->scheme, o->limiter, ijp, ijn, xf, yf, zf, fxp, f->u, f->dUdxi);
    ve = face_value(g, o->scheme, o->limiter, ijp, ijn, xf, yf, zf, fxp, f->v, f->dVdxi);
    we = face_value(g, o->scheme, o->limiter, ijp, ijn, xf, yf, zf, fxp, f->w, f->dWdxi);
  } else {
    ue = face_value(g, o->scheme, o->limiter, ijn, ijp, xf, yf, zf, fxn, f->u, f->dUdxi);
    ve = face_value(g, o->scheme, o->limiter, ijn, ijp, xf, yf, zf, fxn, f->v, f->dVdxi);
    we = face_value(g, o->scheme, o->limiter, ijn, ijp, xf, yf, zf, fxn, f->w, f->dWdxi);
  }
  double fuhigh = flomass * ue, fvhigh = flomass * ve, fwhigh = flomass * we;
  *sup = -gam * (fuhigh - fuuds) + fdue - fdui;
  *svp = -gam * (fvhigh - fvuds) + fdve - fdvi;
  *swp = -gam * (fwhigh - fwuds) + fdwe - fdwi;
}

/* facefluxuvw_boundary (faceflux_velocity.f90:385-549): only `can` (= cb at the call sites) is used by
 * calcuvw; the explicit sources it also computes are dropped there (calcuvw.f90:231-241, :253-263). */
static double facefluxuvw_boundary_can(const fco_mesh *g, const fco_uvw *x, int ijp, int ijb, double xf, double yf,
                                       double zf, double arx, double ary, double arz, double flomass) {
  double xpn = xf - A1(g->xc, ijp), ypn = yf - A1(g->yc, ijp), zpn = zf - A1(g->zc, ijp);
  double dpn = sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
  double are = sqrt(arx * arx + ary * ary + arz * arz);
  double game = A1(x->vis, ijb);
  double de = game * are / dpn;
  return -de + FCO_MIN2(flomass, 0.0);
}

/* presFaceDivInner with the reference's df(ijp,k) addressing (fieldManipulation.f90:395-445) */
static double pres_face_value(const fco_mesh *g, int ijp, int ijn, double xfc, double yfc, double zfc, double fif,
                              const double *fi, const double *df_flat) {
  double fxn = fif, fxp = 1.0 - fxn;
  double xi = A1(g->xc, ijp) * fxp + A1(g->xc, ijn) * fxn;
  double yi = A1(g->yc, ijp) * fxp + A1(g->yc, ijn) * fxn;
  double zi = A1(g->zc, ijp) * fxp + A1(g->zc, ijn) * fxn;
  /* df(i,k) of a dimension(3,numCells) dummy = flat element (i-1) + 3 (k-1), 0-based */
  double dfxi = df_flat[(size_t)(ijp - 1)] * fxp + df_flat[(size_t)(ijn - 1)] * fxn;
  double dfyi = df_flat[(size_t)(ijp - 1) + 3] * fxp + df_flat[(size_t)(ijn - 1) + 3] * fxn;
  double dfzi = df_flat[(size_t)(ijp - 1) + 6] * fxp + df_flat[(size_t)(ijn - 1) + 6] * fxn;
  return A1(fi, ijp) * fxp + A1(fi, ijn) * fxn + dfxi * (xfc - xi) + dfyi * (yfc - yi) + dfzi * (zfc - zi);
}

/* the explicit part of calcuvw up to (not including) the per-component diagonal assembly and solves:
 * su, sv, sw, spu, spv, sp and the off-diagonals of `a` (calcuvw.f90:48-383) */
int fco_calcuvw_assemble(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_uvw_opts *o) {
  const int n = g->numCells;
  if (g->noc > 0 || g->npro > 0) return 2; /* serial src semantics without O-C cuts only */
  if (g->numInnerFaces > 0 && n < 3) return 3; /* df(ijp,3) would leave the gradient array (see header) */
  const int iInletStart = g->numCells + g->npro;
  const int iOutletStart = iInletStart + g->ninl, iSymmetryStart = iOutletStart + g->nout;
  const int iWallStart = iSymmetryStart + g->nsym, iPressOutletStart = iWallStart + g->nwal;
  for (int i = 0; i < n; ++i) { f->su[i] = 0.0; x->sv[i] = 0.0; x->sw[i] = 0.0; x->spu[i] = 0.0; x->spv[i] = 0.0; x->sp[i] = 0.0; }
  fco_grad(g, m, f->u, o->nigrad, f->dUdxi); /* :59-61 */
  fco_grad(g, m, f->v, o->nigrad, f->dVdxi);
  fco_grad(g, m, f->w, o->nigrad, f->dWdxi);
  /* calcPressDiv (fieldManipulation.f90:57-165) */
  for (int istage = 1; istage <= o->nipgrad; ++istage) {
    fco_bpres(g, f->p, f->dPdxi, istage);
    fco_grad(g, m, f->p, o->nigrad, f->dPdxi);
  }
  for (int i = 1; i <= g->numInnerFaces; ++i) {
    int ijp = A1(g->owner, i), ijn = A1(g->neighbour, i);
    double fie = pres_face_value(g, ijp, ijn, A1(g->xf, i), A1(g->yf, i), A1(g->zf, i), A1(g->facint, i), f->p, f->dPdxi);
    double dfxe = fie * A1(g->arx, i), dfye = fie * A1(g->ary, i), dfze = fie * A1(g->arz, i);
    A1(f->su, ijp) = A1(f->su, ijp) - dfxe; A1(x->sv, ijp) = A1(x->sv, ijp) - dfye; A1(x->sw, ijp) = A1(x->sw, ijp) - dfze;
    A1(f->su, ijn) = A1(f->su, ijn) + dfxe; A1(x->sv, ijn) = A1(x->sv, ijn) + dfye; A1(x->sw, ijn) = A1(x->sw, ijn) + dfze;
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
  /* volume sources (calcuvw.f90:75-141) */
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
      if (o->btime > (double)0.99f) { /* `btime > 0.99`: default-real literal */
        sut = sut - apotime * (0.5 * o->btime * A1(x->uoo, inp));
        svt = svt - apotime * (0.5 * o->btime * A1(x->voo, inp));
        swt = swt - apotime * (0.5 * o->btime * A1(x->woo, inp));
      }
      A1(f->su, inp) = A1(f->su, inp) + sut;
      A1(x->sv, inp) = A1(x->sv, inp) + svt;
      A1(x->sw, inp) = A1(x->sw, inp) + swt;
      A1(x->spu, inp) = A1(x->spu, inp) + apotime * (1 + 0.5 * o->btime);
      A1(x->spv, inp) = A1(x->spv, inp) + apotime * (1 + 0.5 * o->btime);
      A1(x->sp, inp) = A1(x->sp, inp) + apotime * (1 + 0.5 * o->btime);
    }
  }
  /* inner faces (:156-186) */
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
  /* inlet, outlet (:225-265) */
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
  /* symmetry (:269-312), wall (:315-383); srds / srdw = are / ((xf-xc).n) (init.f90:987-1029) */
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
  if (o->cn) /* :386-389: the whole array, stale diagonal included */
    for (int k = 0; k < m->nnz; ++k) f->a[k] = 0.5 * f->a[k];
  return 0;
}

/* one velocity component: Crank-Nicolson sources, diagonal, under-relaxation, ap*, BiCGStab
 * (calcuvw.f90:391-441 for U; :447-498 V; :504-556 W).  comp = 0, 1, 2. */
int fco_calcuvw_component(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_uvw_opts *o,
                          int comp, fco_report *rep) {
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
    for (int ijp = 1; ijp <= n; ++ijp) {
      double apotime = A1(f->den, ijp) * A1(g->vol, ijp) / o->timestep;
      double sum = 0.0;
      for (int k = A1(m->ioffset, ijp); k <= A1(m->ioffset, ijp + 1) - 1; ++k) sum = sum + A1(f->a, k);
      double off = sum - A1(f->a, A1(m->diag, ijp));
      A1(s, ijp) = A1(s, ijp) + (apotime + off) * A1(phio, ijp);
      A1(spc, ijp) = A1(spc, ijp) + apotime;
    }
  }
  const double urfrs = 1.0 / o->urf[comp], urfms = 1.0 - o->urf[comp]; /* init.f90:80-81 */
  if (comp > 0)
    for (int inp = 1; inp <= n; ++inp) { A1(f->a, A1(m->diag, inp)) = 0.0; A1(f->su, inp) = 0.0; }
  for (int inp = 1; inp <= n; ++inp) {
    double sum = 0.0;
    for (int k = A1(m->ioffset, inp); k <= A1(m->ioffset, inp + 1) - 1; ++k) sum = sum + A1(f->a, k);
    double off = sum - A1(f->a, A1(m->diag, inp));
    double d = A1(spc, inp) - off;
    d = d * urfrs;
    A1(f->a, A1(m->diag, inp)) = d;
    A1(f->su, inp) = A1(s, inp) + urfms * d * A1(phi, inp);
    A1(ap, inp) = 1.0 / (d + o->sol.small);
  }
  fco_solver_opts so = o->sol;
  so.sor = o->sor[comp];
  so.nsw = o->nsw[comp];
  return fco_bicgstab(m, f->a, f->su, phi, f->res, 0, &so, rep, 0);
}

int fco_calcuvw(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_uvw_opts *o,
                fco_uvw_report *rep) {
  int rc = fco_calcuvw_assemble(g, m, f, x, o);
  if (rc) return rc;
  for (int comp = 0; comp < 3; ++comp) {
    rc = fco_calcuvw_component(g, m, f, x, o, comp, &rep->rep[comp]);
    if (rc) return rc;
  }
  return 0;
}
