/* bsim3_fields.h -- what BSIM3setup/BSIM3temp leave behind for BSIM3load, as flat tables.
 * X(name) names member `BSIM3<name>` of BSIM3model (bsim3def.h:380-...), bsim3SizeDependParam
 * (:263-378) or BSIM3instance (:19-221).  The oracle dump (oracle/ref_hooks.c) and the device
 * kernel share these lists, so the order here IS the table layout. */
#ifndef NGB_BSIM3_FIELDS_H
#define NGB_BSIM3_FIELDS_H

#define NGB_B3_MODEL_FIELDS(X) \
  X(type) X(mobMod) X(capMod) X(acmMod) X(xpart) X(vtm) X(jctEmissionCoeff) X(jctTempSatCurDensity) \
  X(jctSidewallTempSatCurDensity) X(ijth) X(factor1) X(tnom) X(tox) X(cox) X(vcrit) \
  X(unitAreaTempJctCap) X(unitLengthSidewallTempJctCap) X(unitLengthGateSidewallTempJctCap) \
  X(bulkJctBotGradingCoeff) X(bulkJctSideGradingCoeff) X(bulkJctGateSideGradingCoeff) \
  X(PhiB) X(PhiBSW) X(PhiBSWG)

#define NGB_B3_BIN_FIELDS(X) \
  X(k1ox) X(phi) X(weffCV) X(cgbo) X(weff) X(leffCV) X(ckappa) X(cgso) X(cgdo) X(abulkCVfactor) X(uc) \
  X(sqrtPhi) X(ub) X(rds0) X(ldeb) X(ua) X(cgsl) X(cgdl) X(a2) X(a1) X(vbsc) X(pscbe2) X(prwg) X(prwb) \
  X(pdiblb) X(noff) X(ngate) X(litl) X(keta) X(dvt2w) X(dvt2) X(delta) X(cdscd) X(cdscb) X(beta0) \
  X(vsattemp) X(voff) X(thetaRout) X(theta0vb0) X(pscbe1) X(phis3) X(pclm) X(moin) X(kt2) X(k3b) X(k2ox) \
  X(etab) X(dwg) X(dwb) X(dvt0w) X(dvt0) X(cdep0) X(acde) X(a0) X(Xdep0) X(xj) X(w0) X(voffcv) X(vfbcv) \
  X(vbi) X(pvag) X(nlx) X(nfactor) X(leff) X(kt1l) X(kt1) X(k3) X(k1) X(eta0) X(dvt1w) X(dvt1) X(cit) \
  X(cdsc) X(b1) X(b0) X(alpha1) X(alpha0) X(ags)

#define NGB_B3_INST_FIELDS(X) \
  X(m) X(icVDS) X(icVGS) X(icVBS) X(vth0) X(vfb) X(vfbzb) X(u0temp) X(sourceArea) X(drainArea) \
  X(sourcePerimeter) X(drainPerimeter) X(vjsm) X(IsEvjsm) X(vjdm) X(IsEvjdm) X(sourceConductance) \
  X(drainConductance)

enum {
#define X(n) B3M_##n,
  NGB_B3_MODEL_FIELDS(X)
#undef X
  B3M_COUNT
};
enum {
#define X(n) B3P_##n,
  NGB_B3_BIN_FIELDS(X)
#undef X
  B3P_COUNT
};
enum {
#define X(n) B3I_##n,
  NGB_B3_INST_FIELDS(X)
#undef X
  B3I_COUNT
};

/* node roles */
enum { B3N_d, B3N_g, B3N_s, B3N_b, B3N_dp, B3N_sp, B3N_COUNT };

/* flags */
#define B3F_OFF   0x1
#define B3F_NQS   0x2      /* nqsMod or acnqsMod: not on this path (E_UNSUPP) */

/* states, bsim3def.h:223-245 */
enum { B3ST_vbd, B3ST_vbs, B3ST_vgs, B3ST_vds, B3ST_qb, B3ST_cqb, B3ST_qg, B3ST_cqg, B3ST_qd, B3ST_cqd,
       B3ST_qbs, B3ST_qbd, B3ST_qcheq, B3ST_cqcheq, B3ST_qcdump, B3ST_cqcdump, B3ST_qdef, B3ST_COUNT };

/* stamp rows in the statement order of b3ld.c:2981-3069: four rhs adds, then 22 matrix adds */
#define NGB_B3_STAMPS(X) \
  X(rG, g, g) X(rB, b, b) X(rDP, dp, dp) X(rSP, sp, sp) \
  X(Dd, d, d) X(Gg, g, g) X(Ss, s, s) X(Bb, b, b) X(DPdp, dp, dp) X(SPsp, sp, sp) X(Ddp, d, dp) X(Gb, g, b) \
  X(Gdp, g, dp) X(Gsp, g, sp) X(Ssp, s, sp) X(Bg, b, g) X(Bdp, b, dp) X(Bsp, b, sp) X(DPd, dp, d) \
  X(DPg, dp, g) X(DPb, dp, b) X(DPsp, dp, sp) X(SPg, sp, g) X(SPs, sp, s) X(SPb, sp, b) X(SPdp, sp, dp)
enum {
#define X(n, r, c) B3S_##n,
  NGB_B3_STAMPS(X)
#undef X
  B3S_COUNT
};
#define B3S_RHS_COUNT 4
#endif
