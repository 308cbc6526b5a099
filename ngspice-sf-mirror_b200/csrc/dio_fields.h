/* dio_fields.h -- what DIOsetup/DIOtemp leave behind for DIOload, as flat per-instance arrays.
 * X(name) names the member `DIO<name>` of DIOinstance (dio/diodefs.h:30-258); XM(name) the member
 * of DIOmodel (:333-450), copied per instance.  The oracle's dump (oracle/ref_hooks.c) and the
 * device kernel share these lists, so the order here IS the table layout. */
#ifndef NGB_DIO_FIELDS_H
#define NGB_DIO_FIELDS_H

#define NGB_DIO_INST_FIELDS(X) \
  X(temp) X(initCond) X(tVcrit) X(tBrkdwnV) X(tSatCur) X(tSatCur_dT) X(tSatSWCur) X(tSatSWCur_dT) \
  X(tConductance) X(tJctCap) X(tJctPot) X(tDepCap) X(tGradingCoeff) X(tF1) X(tF2) X(tF3) \
  X(tJctSWCap) X(tJctSWPot) X(tDepSWCap) X(tF2SW) X(tF3SW) X(tTransitTime) \
  X(tTunSatCur) X(tTunSatCur_dT) X(tTunSatSWCur) X(tTunSatSWCur_dT) \
  X(forwardKneeCurrent) X(reverseKneeCurrent) X(forwardSWKneeCurrent) X(cmetal) X(cpoly) \
  X(tRecSatCur) X(tRecSatCur_dT) X(tVcritSW) X(tConductanceSW)
#define NGB_DIO_MODEL_FIELDS(X) \
  X(emissionCoeff) X(swEmissionCoeff) X(brkdEmissionCoeff) X(tunEmissionCoeff) X(gradingSWCoeff) X(recEmissionCoeff)

/* what DIOtempUpdate (diotemp.c:18-270) needs to map the parameters to another temperature: read only by instances with
 * self-heating, whose load re-runs it at DIOtemp + delTemp every iteration (dioload.c:317-319).  Appended after the lists
 * above so that fixtures recorded before they existed stay readable (the rows are zero-filled for them) */
#define NGB_DIO_RAW_INST_FIELDS(X) \
  X(tConductance_dT) X(tConductanceSW_dT) X(junctionCap) X(junctionSWCap) X(area) X(m) X(pj)
#define NGB_DIO_RAW_MODEL_FIELDS(X) \
  X(nomTemp) X(gradCoeffTemp1) X(gradCoeffTemp2) X(gradingCoeff) X(tlev) X(tlevc) X(activationEnergy) X(firstBGcorrFactor) \
  X(secndBGcorrFactor) X(junctionPot) X(junctionSWPot) X(tpb) X(tphp) X(cta) X(ctp) X(satCur) X(satSWCur) X(saturationCurrentExp) \
  X(tunSatCur) X(tunSatSWCur) X(tunSaturationCurrentExp) X(tunEGcorrectionFactor) X(recSatCur) X(depletionCapCoeff) \
  X(depletionSWcapCoeff) X(breakdownVoltage) X(tcv) X(level) X(breakdownCurrent) X(tranTimeTemp1) X(tranTimeTemp2) X(transitTime) \
  X(conductance) X(resistTemp1) X(resistTemp2) X(resist) X(conductanceSW) X(resistSW) X(rth0) X(cth0) X(softRevRecParam)

enum {
#define X(n) DIOP_##n,
  NGB_DIO_INST_FIELDS(X)
  NGB_DIO_MODEL_FIELDS(X)
  DIOP_COUNT_V1,                 /* table height of fixtures recorded before the raw fields */
  DIOP_raw_ = DIOP_COUNT_V1 - 1,
  NGB_DIO_RAW_INST_FIELDS(X)
  NGB_DIO_RAW_MODEL_FIELDS(X)
#undef X
  DIOP_COUNT
};

/* flags: instance `off` and the model's Given bits DIOload branches on */
#define DIOF_OFF        0x0001
#define DIOF_BV         0x0002   /* breakdownVoltageGiven      */
#define DIOF_SATSW      0x0004   /* satSWCurGiven              */
#define DIOF_NSW        0x0008   /* swEmissionCoeffGiven       */
#define DIOF_TUNSW      0x0010   /* tunSatSWCurGiven           */
#define DIOF_TUN        0x0020   /* tunSatCurGiven             */
#define DIOF_IKF        0x0040   /* forwardKneeCurrentGiven    */
#define DIOF_IKR        0x0080   /* reverseKneeCurrentGiven    */
#define DIOF_IKP        0x0100   /* forwardSWKneeCurrentGiven  */
#define DIOF_RECSAT     0x0200   /* recSatCurGiven: recombination current with pow-based generation factor */
#define DIOF_RESISTSW   0x0400   /* resistSWGiven: separate sidewall diode behind its own series resistance */
#define DIOF_SELFHEAT   0x0800   /* temperature node `dt` + thermal + rth0 given (dioload.c:80)   */
#define DIOF_REVREC     0x1000   /* soft reverse recovery: qp node, vp != 0, tt != 0 (dioload.c:81) */
#define DIOF_RESIST     0x2000   /* resistGiven (series resistance temperature coefficients apply) */

/* states, diodefs.h:263-289 */
enum { DIOST_voltage, DIOST_current, DIOST_conduct, DIOST_voltageSW, DIOST_currentSW, DIOST_conductSW,
       DIOST_capCharge, DIOST_capCurrent, DIOST_capChargeSW, DIOST_capCurrentSW, DIOST_qth, DIOST_cqth,
       DIOST_deltemp, DIOST_dIdio_dT, DIOST_dIdioSW_dT, DIOST_srcapCharge, DIOST_srcapCurrent, DIOST_qp,
       DIOST_resCurrent, DIOST_resConduct, DIOST_cqcsr, DIOST_gqcsr, DIOST_COUNT };

/* stamp rows in the statement order of dioload.c:779-862: rhs adds (main, sidewall), matrix adds (main, sidewall), the
 * soft-recovery subcircuit; the th* rows exist only with self-heating, the rr* rows only with soft recovery */
enum { DIOS_rhsNeg, DIOS_rhsPosPrime, DIOS_thRhsPos, DIOS_thRhsPp, DIOS_thRhsNeg, DIOS_thRhsTemp,
       DIOS_rhsNegSw, DIOS_rhsPosSwPrime, DIOS_thRhsPosSw, DIOS_thRhsPsp, DIOS_thRhsNegSw, DIOS_thRhsTempSw,
       DIOS_posPos, DIOS_negNeg, DIOS_ppPp, DIOS_posPp, DIOS_negPp, DIOS_ppPos, DIOS_ppNeg,
       DIOS_thTempPos, DIOS_thTempPp, DIOS_thTempNeg, DIOS_thTempTemp, DIOS_thPosTemp, DIOS_thPpTemp, DIOS_thNegTemp,
       DIOS_posPosSw, DIOS_negNegSw, DIOS_pspPsp, DIOS_posPsp, DIOS_negPsp, DIOS_pspPos, DIOS_pspNeg,      /* separate sidewall diode */
       DIOS_thTempPosSw, DIOS_thTempPsp, DIOS_thTempNegSw, DIOS_thPosTempSw, DIOS_thPspTemp, DIOS_thNegTempSw,
       DIOS_rrRhsQp, DIOS_rrQpQp, DIOS_rrQpPp, DIOS_rrQpNeg, DIOS_rrRhsPp, DIOS_rrRhsNeg, DIOS_rrPpQp, DIOS_rrNegQp,
       DIOS_COUNT };
#define DION_COUNT 6              /* node roles: pos, neg, posPrime, posSwPrime, temp, qp */
#endif
