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

enum {
#define X(n) DIOP_##n,
  NGB_DIO_INST_FIELDS(X)
  NGB_DIO_MODEL_FIELDS(X)
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
/* options this path does not implement: reported as E_UNSUPP when the table is added */
#define DIOF_SELFHEAT   0x0800   /* temperature node + rth0                                 */
#define DIOF_REVREC     0x1000   /* soft reverse recovery (qp node)                         */
#define DIOF_UNSUPPORTED (DIOF_SELFHEAT | DIOF_REVREC)

/* states, diodefs.h:263-289 */
enum { DIOST_voltage, DIOST_current, DIOST_conduct, DIOST_voltageSW, DIOST_currentSW, DIOST_conductSW,
       DIOST_capCharge, DIOST_capCurrent, DIOST_capChargeSW, DIOST_capCurrentSW, DIOST_qth, DIOST_cqth,
       DIOST_deltemp, DIOST_dIdio_dT, DIOST_dIdioSW_dT, DIOST_srcapCharge, DIOST_srcapCurrent, DIOST_qp,
       DIOST_resCurrent, DIOST_resConduct, DIOST_cqcsr, DIOST_gqcsr, DIOST_COUNT };

/* stamp rows in the statement order of dioload.c:757-810: rhs adds (main, sidewall), matrix adds (main, sidewall) */
enum { DIOS_rhsNeg, DIOS_rhsPosPrime, DIOS_rhsNegSw, DIOS_rhsPosSwPrime,
       DIOS_posPos, DIOS_negNeg, DIOS_ppPp, DIOS_posPp, DIOS_negPp, DIOS_ppPos, DIOS_ppNeg,
       DIOS_posPosSw, DIOS_negNegSw, DIOS_pspPsp, DIOS_posPsp, DIOS_negPsp, DIOS_pspPos, DIOS_pspNeg,      /* separate sidewall diode */
       DIOS_COUNT };
#endif
