/* vbic_types.h -- tables of the VBIC (4-terminal, no self-heating, no excess phase) load.
 *
 * par  [VBIC_NP][T]   the parameter vector p[] exactly as VBICload assembles it per instance
 *                     (vbic/vbicload.c:127-166: the model block starting at VBICtnom, with the
 *                     temperature-updated instance values written over it)
 * aux  [VBA_COUNT][T] type, tVcrit, icVBE, icVCE, SCALE = area*m, temp
 * nodes [VBN_COUNT][ninst]; states as in vbicdefs.h:440-535 (86 per instance)
 * Self-heating (thermal node + RTH) and excess phase (TD > 0) are refused at upload. */
#ifndef NGB_VBIC_TYPES_H
#define NGB_VBIC_TYPES_H
#include "ngb_types.h"

#define VBIC_NP 108
enum { VBA_type, VBA_tVcrit, VBA_icVBE, VBA_icVCE, VBA_scale, VBA_temp, VBA_COUNT };
enum { VBN_coll, VBN_base, VBN_emit, VBN_subs, VBN_cx, VBN_ci, VBN_bx, VBN_bi, VBN_ei, VBN_bp, VBN_si, VBN_COUNT };
#define VBF_OFF       0x1
#define VBF_SELFHEAT  0x2
#define VBF_EXCESS    0x4
#define VBF_UNSUPPORTED (VBF_SELFHEAT | VBF_EXCESS)

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
  T(si,bx) T(ei,bi) T(bp,bi) T(si,ci) T(si,bi) T(si,bp)

/* the stamps of VBICload in statement order (vbicload.c:1076-1267): R(node, value) adds to the right-hand
 * side, M(row, column, value) to the matrix.  The value expressions name locals of vbic_load_thread. */
#define NGB_VBIC_STAMPS(R, M) \
  R(base, -rc_beo) R(emit, rc_beo) M(base,base, gqbeo) M(emit,emit, gqbeo) M(base,emit, -gqbeo) M(emit,base, -gqbeo) \
  R(base, -rc_bco) R(coll, rc_bco) M(base,base, gqbco) M(coll,coll, gqbco) M(base,coll, -gqbco) M(coll,base, -gqbco) \
  R(bi, -cbcx) R(cx, cbcx) M(bi,bi, gbcx) M(cx,cx, gbcx) M(bi,cx, -gbcx) M(cx,bi, -gbcx) \
  R(bi, -rc_be) M(bi,bi, Ibe_Vbei) M(bi,ei, -Ibe_Vbei) R(ei, rc_be) M(ei,bi, -Ibe_Vbei) M(ei,ei, Ibe_Vbei) \
  R(bx, -rc_bex) M(bx,bx, Ibex_Vbex) M(bx,ei, -Ibex_Vbex) R(ei, rc_bex) M(ei,bx, -Ibex_Vbex) M(ei,ei, Ibex_Vbex) \
  R(ci, -rc_ciei) M(ci,bi, Iciei_Vbei) M(ci,ei, -Iciei_Vbei) M(ci,bi, Iciei_Vbci) M(ci,ci, -Iciei_Vbci) \
  R(ei, rc_ciei) M(ei,bi, -Iciei_Vbei) M(ei,ei, Iciei_Vbei) M(ei,bi, -Iciei_Vbci) M(ei,ci, Iciei_Vbci) \
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
  M(subs,subs, Irs_Vrs) M(si,si, Irs_Vrs) M(si,subs, -Irs_Vrs) M(subs,si, -Irs_Vrs)

/* number of stamp statements */
#define NGB_VBIC_CNT_R(n, v) +1
#define NGB_VBIC_CNT_M(r, c, v) +1
#define VBIC_NSTAMPS (0 NGB_VBIC_STAMPS(NGB_VBIC_CNT_R, NGB_VBIC_CNT_M))

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
