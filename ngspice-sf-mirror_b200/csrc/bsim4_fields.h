/* bsim4_fields.h -- the BSIM4 quantities the load kernel reads, as X-macro lists.
 *
 * One list per storage class.  The same lists generate: the enum indices used by the
 * kernel (B4M_*, B4P_*, B4I_*), the upload tables of the C-ABI, the flattening code of
 * the reference-side shim and the oracle-side dumper (oracle/ref_hooks.c),
 * and the Python fixture loader.  Names follow the reference struct members with the
 * "BSIM4" prefix removed (src/spicelib/devices/bsim4/bsim4def.h: instance :34-669,
 * bsim4SizeDependParam :671-910, model :913-2831).
 */
#ifndef NGB_BSIM4_FIELDS_H
#define NGB_BSIM4_FIELDS_H

/* per-model values (shared by every instance of a .model card; per sample in a Monte-Carlo
 * batch with model-parameter mismatch).  Integer-valued selectors are stored as doubles. */
#define NGB_B4_MODEL_FIELDS(X) \
  X(type) X(mobMod) X(capMod) X(cvchargeMod) X(dioMod) X(rdsMod) X(igcMod) X(igbMod) \
  X(gidlMod) X(tempMod) X(mtrlMod) X(mtrlCompatMod) X(xpart) X(coxe) X(toxe) X(eot) \
  X(epsrox) X(epsrsub) X(epsrgate) X(easub) X(Eg0) X(phig) X(ados) X(bdos) X(factor1) \
  X(vtm) X(vtm0) X(tnom) X(vcrit) X(vtl) X(vtlGiven) X(lambda) X(lambdaGiven) \
  X(pigcdGiven) X(pditsl) X(gidlclamp) X(bvs) X(bvd) X(xjbvs) X(xjbvd) \
  X(vtss) X(vtsd) X(vtssws) X(vtsswd) X(vtsswgs) X(vtsswgd) \
  X(njtsstemp) X(njtsdtemp) X(njtsswstemp) X(njtsswdtemp) X(njtsswgstemp) X(njtsswgdtemp) \
  X(SjctEmissionCoeff) X(DjctEmissionCoeff) \
  X(SjctTempSatCurDensity) X(DjctTempSatCurDensity) \
  X(SjctSidewallTempSatCurDensity) X(DjctSidewallTempSatCurDensity) \
  X(SjctGateSidewallTempSatCurDensity) X(DjctGateSidewallTempSatCurDensity) \
  X(SunitAreaTempJctCap) X(DunitAreaTempJctCap) \
  X(SunitLengthSidewallTempJctCap) X(DunitLengthSidewallTempJctCap) \
  X(SunitLengthGateSidewallTempJctCap) X(DunitLengthGateSidewallTempJctCap) \
  X(SbulkJctBotGradingCoeff) X(DbulkJctBotGradingCoeff) \
  X(SbulkJctSideGradingCoeff) X(DbulkJctSideGradingCoeff) \
  X(SbulkJctGateSideGradingCoeff) X(DbulkJctGateSideGradingCoeff) \
  X(PhiBS) X(PhiBD) X(PhiBSWS) X(PhiBSWD) X(PhiBSWGS) X(PhiBSWGD)

/* per-(L,W,NF)-bin values (pParam->) */
#define NGB_B4_BIN_FIELDS(X) \
  X(Aechvb) X(AechvbEdgeD) X(AechvbEdgeS) X(Bechvb) X(BechvbEdge) X(ToxRatio) X(VgsteffVth) \
  X(Xdep0) X(a0) X(a1) X(a2) X(abulkCVfactor) X(acde) X(agidl) X(agisl) X(ags) X(aigbacc) \
  X(aigbinv) X(aigc) X(aigd) X(aigs) X(alpha0) X(alpha1) X(b0) X(b1) X(beta0) X(bgidl) \
  X(bgisl) X(bigbacc) X(bigbinv) X(bigc) X(bigd) X(bigs) X(cdep0) X(cdsc) X(cdscb) X(cdscd) \
  X(cgbo) X(cgdl) X(cgdo) X(cgidl) X(cgisl) X(cgsl) X(cgso) X(cigbacc) X(cigbinv) X(cigc) \
  X(cigd) X(cigs) X(cit) X(ckappad) X(ckappas) X(delta) X(dvt0) X(dvt0w) X(dvt1) X(dvt1w) \
  X(dvt2) X(dvt2w) X(dvtp0) X(dvtp1) X(dvtp2factor) X(dvtp4) X(dwb) X(dwg) X(egidl) X(egisl) \
  X(eigbinv) X(etab) X(eu) X(fgidl) X(fgisl) X(fprout) X(k1) X(k1ox) X(k3) X(k3b) X(keta) \
  X(ketac) X(kgidl) X(kgisl) X(kt1) X(kt1l) X(kt2) X(lambda) X(ldeb) X(leff) X(leffCV) \
  X(litl) X(lpe0) X(lpeb) X(moin) X(mstar) X(mstarcv) X(nfactor) X(ngate) X(nigbacc) \
  X(nigbinv) X(nigc) X(noff) X(pclm) X(pdiblb) X(pdits) X(pditsd) X(phi) X(pigcd) X(prwb) \
  X(prwg) X(pscbe1) X(pscbe2) X(pvag) X(rd0) X(rds0) X(rdswmin) X(rdwmin) X(rgidl) X(rgisl) \
  X(rs0) X(rswmin) X(sqrtPhi) X(tfactor) X(theta0vb0) X(thetaRout) X(ua) X(ub) X(uc) X(ucs) \
  X(ud) X(vbi) X(vfbcv) X(vfbsd) X(vfbsdoff) X(voffcbn) X(voffcbncv) X(voffcv) X(vtl) X(w0) \
  X(weff) X(weffCJ) X(weffCV) X(xj) X(xrcrg1) X(xrcrg2)

/* per-instance read-only values produced by BSIM4setup/BSIM4temp (here->) */
#define NGB_B4_INST_FIELDS(X) \
  X(Adeff) X(Aseff) X(Pdeff) X(Pseff) X(nf) X(m) X(mult_i) X(mult_q) \
  X(icVDS) X(icVGS) X(icVBS) X(vth0) X(vfb) X(vfbzb) X(vbsc) X(k2ox) X(eta0) X(toxp) X(coxp) \
  X(u0temp) X(vsattemp) X(vtfbphi1) X(vtfbphi2) X(grgeltd) X(grbdb) X(grbpb) X(grbpd) \
  X(grbps) X(grbsb) X(sourceConductance) X(drainConductance) \
  X(vjsmFwd) X(vjsmRev) X(vjdmFwd) X(vjdmRev) X(IVjsmFwd) X(IVjsmRev) X(IVjdmFwd) X(IVjdmRev) \
  X(SslpFwd) X(SslpRev) X(DslpFwd) X(DslpRev) X(XExpBVS) X(XExpBVD) \
  X(SjctTempRevSatCur) X(DjctTempRevSatCur) X(SswTempRevSatCur) X(DswTempRevSatCur) \
  X(SswgTempRevSatCur) X(DswgTempRevSatCur)

/* per-instance integer selectors, packed by the host into one int32 (see B4F_* below) */
#define NGB_B4_INST_INT_FIELDS(X) X(off) X(rbodyMod) X(rgateMod) X(trnqsMod) X(acnqsMod)

/* node roles, in the order node indices are uploaded ([12][ninst]); bsim4def.h:44-55 */
#define NGB_B4_NODE_FIELDS(X) \
  X(dNode) X(gNodeExt) X(sNode) X(bNode) X(dNodePrime) X(gNodePrime) X(gNodeMid) \
  X(sNodePrime) X(bNodePrime) X(dbNode) X(sbNode) X(qNode)

/* matrix stamp positions (the 70 TSTALLOC'd pointers, b4set.c:2587-2676, bsim4def.h:332-413)
 * as (row role, column role).  Order is ours; the value each receives per evaluation is the
 * sum of every `+=`/`-=` the reference applies to that pointer (b4ld.c:5235-5388). */
#define NGB_B4_MAT_FIELDS(X) \
  X(GEge) X(GPge) X(GEgp) X(GPgp) X(GPdp) X(GPsp) X(GPbp) X(GEdp) X(GEsp) X(GEbp) \
  X(GEgm) X(GMge) X(GMgm) X(GMdp) X(GMgp) X(GMsp) X(GMbp) X(DPgm) X(GPgm) X(SPgm) X(BPgm) \
  X(Dgp) X(Dsp) X(Dbp) X(Sdp) X(Sgp) X(Sbp) \
  X(DPdp) X(DPd) X(DPgp) X(DPsp) X(DPbp) X(Ddp) X(Dd) \
  X(SPdp) X(SPgp) X(SPsp) X(SPs) X(SPbp) X(Ssp) X(Ss) \
  X(BPdp) X(BPgp) X(BPsp) X(BPbp) \
  X(DPdb) X(SPsb) X(DBdp) X(DBdb) X(DBbp) X(DBb) X(BPdb) X(BPb) X(BPsb) \
  X(SBsp) X(SBbp) X(SBb) X(SBsb) X(Bdb) X(Bbp) X(Bsb) X(Bb) \
  X(Qq) X(Qgp) X(Qdp) X(Qsp) X(Qbp) X(DPq) X(SPq) X(GPq)

/* right-hand-side stamp positions (node roles that receive current, b4ld.c:5024-5053) */
#define NGB_B4_RHS_FIELDS(X) \
  X(dp) X(gp) X(ge) X(gm) X(bp) X(sp) X(db) X(sb) X(d) X(s) X(q)

enum {
#define X(n) B4M_##n,
  NGB_B4_MODEL_FIELDS(X)
#undef X
  B4M_COUNT
};
enum {
#define X(n) B4P_##n,
  NGB_B4_BIN_FIELDS(X)
#undef X
  B4P_COUNT
};
enum {
#define X(n) B4I_##n,
  NGB_B4_INST_FIELDS(X)
#undef X
  B4I_COUNT
};
enum {
#define X(n) B4N_##n,
  NGB_B4_NODE_FIELDS(X)
#undef X
  B4N_COUNT
};
enum {
#define X(n) B4S_##n,
  NGB_B4_MAT_FIELDS(X)
#undef X
  B4S_MAT_COUNT
};
enum {
  B4R_FIRST_ = B4S_MAT_COUNT - 1,
#define X(n) B4R_##n,
  NGB_B4_RHS_FIELDS(X)
#undef X
  B4S_COUNT               /* matrix + rhs stamp positions per instance */
};

/* "exact-order" extras: the reference applies several separate `+=` to some pointers (channel
 * term, then GIDL, then GISL, then the body network: b4ld.c:5294-5364).  In exact-order mode
 * each addend gets its own stamp row so that Ax is summed in precisely the reference order;
 * in merged mode these positions are unused (-1) and the addends are pre-summed. */
#define NGB_B4_EXTRA_FIELDS(X) \
  X(DPdp_g) X(DPgp_g) X(DPsp_g) X(DPbp_g) X(BPdp_g) X(BPgp_g) X(BPsp_g) X(BPbp_g) \
  X(SPdp_s) X(SPgp_s) X(SPsp_s) X(SPbp_s) X(BPdp_s) X(BPgp_s) X(BPsp_s) X(BPbp_s) X(BPbp_r)
enum {
  B4X_FIRST_ = B4S_COUNT - 1,
#define X(n) B4X_##n,
  NGB_B4_EXTRA_FIELDS(X)
#undef X
  B4S_TOTAL               /* rows of the per-instance stamp-position table */
};

/* packed per-instance flags */
#define B4F_OFF        0x1
#define B4F_RBODY_SH   1      /* 2 bits */
#define B4F_RGATE_SH   3      /* 2 bits */
#define B4F_RBODY(f)   (((f) >> B4F_RBODY_SH) & 3)
#define B4F_RGATE(f)   (((f) >> B4F_RGATE_SH) & 3)

/* state vector layout per instance (bsim4def.h:534-567) */
enum { B4ST_vbd = 0, B4ST_vbs, B4ST_vgs, B4ST_vds, B4ST_vdbs, B4ST_vdbd, B4ST_vsbs, B4ST_vges,
       B4ST_vgms, B4ST_vses, B4ST_vdes, B4ST_qb, B4ST_cqb, B4ST_qg, B4ST_cqg, B4ST_qd, B4ST_cqd,
       B4ST_qgmid, B4ST_cqgmid, B4ST_qbs, B4ST_cqbs, B4ST_qbd, B4ST_cqbd, B4ST_qcheq, B4ST_cqcheq,
       B4ST_qcdump, B4ST_cqcdump, B4ST_qdef, B4ST_qs, B4ST_COUNT };

/* operating-point values kept between Newton iterations / exported for parity checks */
#define NGB_B4_OP_FIELDS(X) \
  X(von) X(mode) X(cd) X(gm) X(gds) X(gmbs) X(gbd) X(gbs) X(cbd) X(cbs) X(csub) X(gbbs) X(gbgs) \
  X(gbds) X(Igidl) X(Igisl) X(Igcs) X(Igcd) X(Igs) X(Igd) X(Igb) X(vdsat) X(Vgsteff) X(Vdseff) \
  X(qgate) X(qbulk) X(qdrn) X(capbd) X(capbs) X(cggb) X(cgdb) X(cgsb) X(cbgb) X(cbdb) X(cbsb) \
  X(cdgb) X(cddb) X(cdsb)
enum {
#define X(n) B4O_##n,
  NGB_B4_OP_FIELDS(X)
#undef X
  B4O_COUNT
};

#endif
