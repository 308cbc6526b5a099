/* vbic_types.h -- tables of the VBIC (4-terminal, electro-thermal, excess phase) load.
 *
 * par  [VBIC_NP][T]   the parameter vector p[] exactly as VBICload assembles it per instance
 *                     (vbic/vbicload.c:127-166: the model block starting at VBICtnom, with the
 *                     temperature-updated instance values written over it)
 * aux  [VBA_COUNT][T] type, tVcrit, icVBE, icVCE, SCALE = area*m, temp
 * nodes [VBN_COUNT][ninst]; states as in vbicdefs.h:440-535 (86 per instance)
 * Self-heating (VBF_SELFHEAT: thermal node `dt`, RTH = p[83], CTH = p[84]) and excess phase (VBF_EXCESS: the
 * second-order filter nodes xf1 / xf2, TD = p[62]) add node roles, structural entries and stamps that exist only for the
 * instances that carry the flag (vbicsetup.c:470-523, 584-624). */
#ifndef NGB_VBIC_TYPES_H
#define NGB_VBIC_TYPES_H
#include "ngb_types.h"

#define VBIC_NP 108
enum { VBA_type, VBA_tVcrit, VBA_icVBE, VBA_icVCE, VBA_scale, VBA_temp, VBA_COUNT };
enum { VBN_coll, VBN_base, VBN_emit, VBN_subs, VBN_cx, VBN_ci, VBN_bx, VBN_bi, VBN_ei, VBN_bp, VBN_si, VBN_temp, VBN_xf1, VBN_xf2, VBN_COUNT };
#define VBF_OFF       0x1
#define VBF_SELFHEAT  0x2
#define VBF_EXCESS    0x4

/* states, vbicdefs.h:440-535 */
enum { VBS_vbei, VBS_vbex, VBS_vbci, VBS_vbcx, VBS_vbep, VBS_vrci, VBS_vrbi, VBS_vrbp, VBS_vbcp, VBS_vxf1, VBS_vxf2, VBS_ibe,
       VBS_ibe_Vbei, VBS_ibex, VBS_ibex_Vbex, VBS_iciei, VBS_iciei_Vbei, VBS_iciei_Vbci, VBS_iciei_Vrth, VBS_iciei_Vxf2,
       VBS_ibc, VBS_ibc_Vbci, VBS_ibc_Vbei, VBS_ibep, VBS_ibep_Vbep, VBS_irci, VBS_irci_Vrci, VBS_irci_Vbci, VBS_irci_Vbcx,
       VBS_irbi, VBS_irbi_Vrbi, VBS_irbi_Vbei, VBS_irbi_Vbci, VBS_irbp, VBS_irbp_Vrbp, VBS_irbp_Vbep, VBS_irbp_Vbci,
       VBS_qbe, VBS_cqbe, VBS_cqbeci, VBS_qbex, VBS_cqbex, VBS_qbc, VBS_cqbc, VBS_qbcx, VBS_cqbcx, VBS_qbep, VBS_cqbep,
       VBS_cqbepci, VBS_qbeo, VBS_cqbeo, VBS_gqbeo, VBS_qbco, VBS_cqbco, VBS_gqbco, VBS_ibcp, VBS_ibcp_Vbcp, VBS_iccp,
       VBS_iccp_Vbep, VBS_iccp_Vbci, VBS_iccp_Vbcp, VBS_qbcp, VBS_cqbcp, VBS_ircx_Vrcx, VBS_irbx_Vrbx, VBS_irs_Vrs, VBS_ire_Vre,
       VBS_qcth, VBS_cqcth, VBS_vrth, VBS_icth_Vrth, VBS_qxf1, VBS_cqxf1, VBS_gqxf1, VBS_ixf1, VBS_ixf1_Vbei, VBS_ixf1_Vbci,
       VBS_ixf1_Vxf1, VBS_ixf1_Vxf2, VBS_ixf1_Vrth, VBS_qxf2, VBS_cqxf2, VBS_gqxf2, VBS_ixf2, VBS_ixf2_Vxf1, VBS_ixf2_Vxf2,
       VBS_COUNT };

/* structural entries requested by VBICsetup (vbicsetup.c:530-582), as (row, column) node roles */
#define NGB_VBIC_STRUCT(T) \
  T(coll,coll) T(base,base) T(emit,emit) T(subs,subs) T(cx,cx) T(ci,ci) T(bx,bx) T(bi,bi) T(ei,ei) T(bp,bp) T(si,si) \
  T(base,emit) T(emit,base) T(base,coll) T(coll,base) T(coll,cx) T(base,bx) T(emit,ei) T(subs,si) T(cx,ci) T(cx,bx) \
  T(cx,bi) T(cx,bp) T(ci,bi) T(ci,ei) T(bx,bi) T(bx,ei) T(bx,bp) T(bx,si) T(bi,ei) T(bp,si) T(cx,coll) T(bx,base) \
  T(ei,emit) T(si,subs) T(ci,cx) T(bi,cx) T(bp,cx) T(bx,ci) T(bi,ci) T(ei,ci) T(bp,ci) T(bi,bx) T(ei,bx) T(bp,bx) \
  T(si,bx) T(ei,bi) T(bp,bi) T(si,ci) T(si,bi) T(si,bp) \
  /* self-heating (vbicsetup.c:584-611): entries with the thermal node; absent (node 0) without it */ \
  T(coll,temp) T(base,temp) T(emit,temp) T(subs,temp) T(ci,temp) T(cx,temp) T(bi,temp) T(bx,temp) T(bp,temp) T(ei,temp) T(si,temp) \
  T(temp,coll) T(temp,ci) T(temp,cx) T(temp,bi) T(temp,base) T(temp,bx) T(temp,bp) T(temp,emit) T(temp,ei) T(temp,subs) T(temp,si) \
  T(temp,temp) T(xf1,temp) \
  /* excess phase (vbicsetup.c:613-624) */ \
  T(xf1,xf1) T(xf1,bi) T(xf1,ei) T(xf1,ci) T(xf1,xf2) T(xf2,xf1) T(xf2,xf2) T(ei,xf2) T(ci,xf2)

/* the stamps of VBICload in statement order (vbicload.c:1076-1475): R(node, value) adds to the right-hand
 * side, M(row, column, value) to the matrix; RX / MX exist only for instances with excess phase, RS / MS only with
  * self-heating, RSX / MSX with both.  The value expressions name locals of vbic_load_thread. */
#define NGB_VBIC_STAMPS(R, M, RX, MX, RS, MS, RSX, MSX) \
  R(base, -rc_beo) R(emit, rc_beo) M(base,base, gqbeo) M(emit,emit, gqbeo) M(base,emit, -gqbeo) M(emit,base, -gqbeo) \
  R(base, -rc_bco) R(coll, rc_bco) M(base,base, gqbco) M(coll,coll, gqbco) M(base,coll, -gqbco) M(coll,base, -gqbco) \
  R(bi, -cbcx) R(cx, cbcx) M(bi,bi, gbcx) M(cx,cx, gbcx) M(bi,cx, -gbcx) M(cx,bi, -gbcx) \
  R(bi, -rc_be) M(bi,bi, Ibe_Vbei) M(bi,ei, -Ibe_Vbei) R(ei, rc_be) M(ei,bi, -Ibe_Vbei) M(ei,ei, Ibe_Vbei) \
  R(bx, -rc_bex) M(bx,bx, Ibex_Vbex) M(bx,ei, -Ibex_Vbex) R(ei, rc_bex) M(ei,bx, -Ibex_Vbex) M(ei,ei, Ibex_Vbex) \
  R(ci, -rc_ciei) M(ci,bi, Iciei_Vbei) M(ci,ei, -Iciei_Vbei) M(ci,bi, Iciei_Vbci) M(ci,ci, -Iciei_Vbci) \
  R(ei, rc_ciei) M(ei,bi, -Iciei_Vbei) M(ei,ei, Iciei_Vbei) M(ei,bi, -Iciei_Vbci) M(ei,ci, Iciei_Vbci) \
  RX(ci, -rc_cixf) RX(ei, rc_cixf) MX(ci,xf2, Iciei_Vxf2) MX(ei,xf2, -Iciei_Vxf2) \
  R(bi, -rc_bc) M(bi,bi, Ibc_Vbci) M(bi,ci, -Ibc_Vbci) M(bi,bi, Ibc_Vbei) M(bi,ei, -Ibc_Vbei) \
  R(ci, rc_bc) M(ci,bi, -Ibc_Vbci) M(ci,ci, Ibc_Vbci) M(ci,bi, -Ibc_Vbei) M(ci,ei, Ibc_Vbei) \
  R(bx, -rc_bep) M(bx,bx, Ibep_Vbep) M(bx,bp, -Ibep_Vbep) R(bp, rc_bep) M(bp,bx, -Ibep_Vbep) M(bp,bp, Ibep_Vbep) \
  M(coll,coll, Ircx_Vrcx) M(cx,cx, Ircx_Vrcx) M(cx,coll, -Ircx_Vrcx) M(coll,cx, -Ircx_Vrcx) \
  R(cx, -rc_rci) M(cx,cx, Irci_Vrci) M(cx,ci, -Irci_Vrci) M(cx,bi, Irci_Vbci) M(cx,ci, -Irci_Vbci) M(cx,bi, Irci_Vbcx) M(cx,cx, -Irci_Vbcx) \
  R(ci, rc_rci) M(ci,cx, -Irci_Vrci) M(ci,ci, Irci_Vrci) M(ci,bi, -Irci_Vbci) M(ci,ci, Irci_Vbci) M(ci,bi, -Irci_Vbcx) M(ci,cx, Irci_Vbcx) \
  M(base,base, Irbx_Vrbx) M(bx,bx, Irbx_Vrbx) M(bx,base, -Irbx_Vrbx) M(base,bx, -Irbx_Vrbx) \
  R(bx, -rc_rbi) M(bx,bx, Irbi_Vrbi) M(bx,bi, -Irbi_Vrbi) M(bx,bi, Irbi_Vbei) M(bx,ei, -Irbi_Vbei) M(bx,bi, Irbi_Vbci) M(bx,ci, -Irbi_Vbci) \
  R(bi, rc_rbi) M(bi,bx, -Irbi_Vrbi) M(bi,bi, Irbi_Vrbi) M(bi,bi, -Irbi_Vbei) M(bi,ei, Irbi_Vbei) M(bi,bi, -Irbi_Vbci) M(bi,ci, Irbi_Vbci) \
  M(emit,emit, Ire_Vre) M(ei,ei, Ire_Vre) M(ei,emit, -Ire_Vre) M(emit,ei, -Ire_Vre) \
  R(bp, -rc_rbp) M(bp,bp, Irbp_Vrbp) M(bp,cx, -Irbp_Vrbp) M(bp,bx, Irbp_Vbep) M(bp,bp, -Irbp_Vbep) M(bp,bi, Irbp_Vbci) M(bp,ci, -Irbp_Vbci) \
  R(cx, rc_rbp) M(cx,bp, -Irbp_Vrbp) M(cx,cx, Irbp_Vrbp) M(cx,bx, -Irbp_Vbep) M(cx,bp, Irbp_Vbep) M(cx,bi, -Irbp_Vbci) M(cx,ci, Irbp_Vbci) \
  R(si, -rc_bcp) M(si,si, Ibcp_Vbcp) M(si,bp, -Ibcp_Vbcp) R(bp, rc_bcp) M(bp,si, -Ibcp_Vbcp) M(bp,bp, Ibcp_Vbcp) \
  R(bx, -rc_ccp) M(bx,bx, Iccp_Vbep) M(bx,bp, -Iccp_Vbep) M(bx,bi, Iccp_Vbci) M(bx,ci, -Iccp_Vbci) M(bx,si, Iccp_Vbcp) M(bx,bp, -Iccp_Vbcp) \
  R(si, rc_ccp) M(si,bx, -Iccp_Vbep) M(si,bp, Iccp_Vbep) M(si,bi, -Iccp_Vbci) M(si,ci, Iccp_Vbci) M(si,si, -Iccp_Vbcp) M(si,bp, Iccp_Vbcp) \
  M(subs,subs, Irs_Vrs) M(si,si, Irs_Vrs) M(si,subs, -Irs_Vrs) M(subs,si, -Irs_Vrs) \
  /* excess phase: the two filter nodes (:1269-1285) */ \
  RX(xf1, -rc_xf1) MX(xf1,bi, Ixf1_Vbei) MX(xf1,ei, -Ixf1_Vbei) MX(xf1,bi, Ixf1_Vbci) MX(xf1,ci, -Ixf1_Vbci) MX(xf1,xf2, Ixf1_Vxf2) MX(xf1,xf1, Ixf1_Vxf1) \
  RX(xf2, -rc_xf2) MX(xf2,xf2, Ixf2_Vxf2) MX(xf2,xf1, Ixf2_Vxf1) \
  /* self-heating: d/dVrth of every branch current (:1287-1388) */ \
  RS(bi, -rt_be) MS(bi,temp, Ibe_Vrth) RS(ei, rt_be) MS(ei,temp, -Ibe_Vrth) \
  RS(bx, -rt_bex) MS(bx,temp, Ibex_Vrth) RS(ei, rt_bex) MS(ei,temp, -Ibex_Vrth) \
  RS(ci, -rt_ciei) MS(ci,temp, Iciei_Vrth) RS(ei, rt_ciei) MS(ei,temp, -Iciei_Vrth) \
  RS(bi, -rt_bc) MS(bi,temp, Ibc_Vrth) RS(ci, rt_bc) MS(ci,temp, -Ibc_Vrth) \
  RS(bx, -rt_bep) MS(bx,temp, Ibep_Vrth) RS(bp, rt_bep) MS(bp,temp, -Ibep_Vrth) \
  RS(coll, -rt_rcx) MS(coll,temp, Ircx_Vrth) RS(cx, rt_rcx) MS(cx,temp, -Ircx_Vrth) \
  RS(cx, -rt_rci) MS(cx,temp, Irci_Vrth) RS(ci, rt_rci) MS(ci,temp, -Irci_Vrth) \
  RS(base, -rt_rbx) MS(base,temp, Irbx_Vrth) RS(bx, rt_rbx) MS(bx,temp, -Irbx_Vrth) \
  RS(bx, -rt_rbi) MS(bx,temp, Irbi_Vrth) RS(bi, rt_rbi) MS(bi,temp, -Irbi_Vrth) \
  RS(emit, -rt_re) MS(emit,temp, Ire_Vrth) RS(ei, rt_re) MS(ei,temp, -Ire_Vrth) \
  RS(bp, -rt_rbp) MS(bp,temp, Irbp_Vrth) RS(cx, rt_rbp) MS(cx,temp, -Irbp_Vrth) \
  RS(si, -rt_bcp) MS(si,temp, Ibcp_Vrth) RS(bp, rt_bcp) MS(bp,temp, -Ibcp_Vrth) \
  RS(bx, -rt_ccp) MS(bx,temp, Iccp_Vrth) RS(si, rt_ccp) MS(si,temp, -Iccp_Vrth) \
  RS(subs, -rt_rs) MS(subs,temp, Irs_Vrth) RS(si, rt_rs) MS(si,temp, -Irs_Vrth) \
  /* thermal network: Rth, Cth, dissipated power (:1389-1462) */ \
  MS(temp,temp, Irth_Vrth) RS(temp, -rc_cth) MS(temp,temp, Icth_Vrth) RS(temp, rc_ith) MS(temp,temp, -Ith_Vrth) \
  MS(temp,bi, -Ith_Vbei) MS(temp,ei, Ith_Vbei) MS(temp,bi, -Ith_Vbci) MS(temp,ci, Ith_Vbci) MS(temp,ci, -Ith_Vcei) MS(temp,ei, Ith_Vcei) \
  MS(temp,bx, -Ith_Vbex) MS(temp,ei, Ith_Vbex) MS(temp,bx, -Ith_Vbep) MS(temp,bp, Ith_Vbep) MS(temp,subs, -Ith_Vbcp) MS(temp,bp, Ith_Vbcp) \
  MS(temp,bx, -Ith_Vcep) MS(temp,subs, Ith_Vcep) MS(temp,cx, -Ith_Vrci) MS(temp,ci, Ith_Vrci) MS(temp,bi, -Ith_Vbcx) MS(temp,cx, Ith_Vbcx) \
  MS(temp,bx, -Ith_Vrbi) MS(temp,bi, Ith_Vrbi) MS(temp,bp, -Ith_Vrbp) MS(temp,cx, Ith_Vrbp) MS(temp,coll, -Ith_Vrcx) MS(temp,cx, Ith_Vrcx) \
  MS(temp,base, -Ith_Vrbx) MS(temp,bx, Ith_Vrbx) MS(temp,emit, -Ith_Vre) MS(temp,ei, Ith_Vre) MS(temp,subs, -Ith_Vrs) MS(temp,si, Ith_Vrs) \
  RSX(xf1, -rt_xf1) MSX(xf1,temp, Ixf1_Vrth)

/* number of stamp statements */
#define NGB_VBIC_CNT_R(n, v) +1
#define NGB_VBIC_CNT_M(r, c, v) +1
#define VBIC_NSTAMPS (0 NGB_VBIC_STAMPS(NGB_VBIC_CNT_R, NGB_VBIC_CNT_M, NGB_VBIC_CNT_R, NGB_VBIC_CNT_M, NGB_VBIC_CNT_R, NGB_VBIC_CNT_M, NGB_VBIC_CNT_R, NGB_VBIC_CNT_M))

typedef struct NgbVbicCtx {
    int ninst, S, T, nstamps;
    const int *nodes;       /* [VBN_COUNT][ninst]                                     */
    const int *flags;       /* [ninst] VBF_*                                           */
    const double *par;      /* [VBIC_NP][T]                                            */
    const double *aux;      /* [VBA_COUNT][T]                                          */
    const int *spos;        /* [nstamps][ninst] stamp rows, -1 = ground                */
    double *state;          /* [nhist][VBS_COUNT][T]                                   */
    double *stamp;
    const double *x; int neq1;
    NgbCtl ctl;
} NgbVbicCtx;
#endif
