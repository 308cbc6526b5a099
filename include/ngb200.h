/* ngb200.h -- C ABI of the B200 Newton hot path (libngb200.so).
 *
 * Plain pointers and sizes only.  Every entry point names the reference interface it stands
 * in for (paths relative to the ngspice tree, danchitnis/ngspice-sf-mirror):
 *
 *   device table   SPICEdev::DEVload / DEVtrunc      src/include/ngspice/devdefs.h:57,67
 *   circuit load   CKTload                            src/spicelib/analysis/cktload.c:32
 *   matrix API     SMPclear/SMPluFac/SMPreorder/SMPsolve/SMPpreOrder
 *                                                     src/include/ngspice/smpdefs.h:61-86
 *   Newton loop    NIiter, NIconvTest                 src/maths/ni/niiter.c:28, niconv.c:21
 *   transient      DCtran, NIcomCof, CKTtrunc/CKTterr src/spicelib/analysis/dctran.c:66,
 *                                                     src/maths/ni/nicomcof.c:14, cktterr.c:10
 *
 * A `ngb_circuit` is the flattened result of CKTsetup + CKTtemp for one netlist (what the
 * reference keeps in CKTcircuit, the GENinstance lists and the KLU binding table).  A
 * `ngb_batch` is S device-resident samples of that circuit (Monte-Carlo draws or sweep
 * points): parameters, solution vectors, state history, matrices.  All array arguments are
 * HOST pointers unless the name says `dev`; layouts are given per call.
 *
 * Return value: 0 (OK) or a code from src/include/ngspice/sperror.h / iferrmsg.h
 * (E_SINGULAR 102, E_ITERLIM 103, E_ORDER 104, E_METHOD 105, E_TIMESTEP 106, E_UNSUPP 10).
 * There is no CPU fallback: every compute entry point fails with E_PANIC when the CUDA
 * device cannot be initialised.
 */
#ifndef NGB200_H
#define NGB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngb_circuit ngb_circuit;
typedef struct ngb_batch ngb_batch;

/* ---- library ---- */
const char *ngbBackend(void);                 /* "cuda-sm_100a" for the product library */
const char *ngbLastError(void);
long ngbLaunchCount(void);                    /* kernels launched since load */
int ngbInit(int device);                      /* select the CUDA device, create the launch stream */
int ngbSetStream(void *cuda_stream);          /* launch on a caller-owned cudaStream_t instead */
int ngbSync(void);
/* CUDA-event timing of the dominant kernel: every `every`-th BSIM4 load launch is bracketed by
 * events; ngbProfileRead returns the summed milliseconds and the number of timed launches */
void ngbProfile(int enable, int every);
int ngbProfileRead(double *ms_sum, long *count);
/* FP64-pipe peak of the selected GPU, measured by a register-only micro-benchmark (the roofline's FP64 denominator):
 * out[0] DFMA flop/s (2 per instruction), out[1] DADD/DMUL flop/s (1 per instruction: what -fmad=false code can reach),
 * out[2] SM clock in kHz as the driver reports it */
int ngbMeasureFp64Peak(double out[3]);

/* field-list sizes, so callers can check they were built against the same lists
 * (bsim4_fields.h): [0]=model [1]=bin [2]=instance [3]=node roles [4]=matrix stamps
 * [5]=matrix+rhs stamps [6]=states [7]=op-point fields; list 8 of ngbBsim4FieldName is the diode
 * parameter list of dio_fields.h, lists 9-11 the BSIM3 model / bin / instance lists (ngbDioLayout: [0]=parameters [1]=states [2]=stamp rows) */
void ngbBsim4Layout(int out[8]);
const char *ngbBsim4FieldName(int list, int index);
void ngbDioLayout(int out[3]);

/* ---- circuit description: what DEVsetup/DEVtemperature/DEVbindCSC leave behind ---- */
ngb_circuit *ngbCircuitCreate(int neq, const int *node_type /* [neq+1], 3=voltage 4=current */);
void ngbCircuitDestroy(ngb_circuit *c);

/* dopt: reltol abstol vntol chgtol trtol temp vt0 xmu tstep tstop tmax tstart delmin minbreak gmin
 * iopt: method(1 trap) maxorder itl4 itl1 uic   (cktntask.c:95-146, cktdojob.c:50-125) */
int ngbCircuitSetOptions(ngb_circuit *c, const double dopt[15], const int iopt[5]);
/* CKTop's fallbacks after a failed plain NIiter (cktop.c:62-96): CKTnumGminSteps / CKTnumSrcSteps (defaults 1 / 1,
 * cktntask.c:120-121; 0 skips the route; 1 = dynamic_gmin then new_gmin / gillespie_src; > 1 would be spice3_gmin /
 * spice3_src: E_UNSUPP), CKTdcTrcvMaxIter (itl2, default 50), CKTgminFactor (default 10) and CKTnoOpIter (`.option noopiter`:
 * CKTop skips the plain NIiter, cktop.c:42-55).  They run per sample inside
 * the device controller of ngbTranRun */
int ngbCircuitSetOpFallbacks(ngb_circuit *c, int num_gmin_steps, int num_src_steps, int itl2, double gmin_factor, int no_op_iter,
                             double gshunt /* CKTgshunt: the ladders end on MAX(CKTgmin, CKTgshunt) and leave CKTdiagGmin = CKTgshunt */);

/* BSIM4 instances in the order of the reference instance lists (cktcrte.c:62-64).
 * nodes [12][ninst], flags [ninst] (B4F_*), prow [ninst] row of mtab/ptab, inst [NI][ninst],
 * mtab [nrows][NM], ptab [nrows][NP]            -- replaces the BSIM4instance/BSIM4model walk
 * of BSIM4load (b4ld.c:251-253) */
int ngbCircuitAddBsim4(ngb_circuit *c, int ninst, const int *nodes, const int *flags, const int *prow,
                       const double *inst, int nrows, const double *mtab, const double *ptab);
/* 1 (default): Ax/rhs are summed in exactly the reference's statement order (one stamp row per
 * `+=` of b4ld.c:5235-5388); 0: addends to one pointer are pre-summed (fewer stamp rows) */
int ngbCircuitSetExactOrder(ngb_circuit *c, int on);
int ngbCircuitAddResistors(ngb_circuit *c, int n, const int *nodes /* [2][n] */, const double *g);
int ngbCircuitAddCapacitors(ngb_circuit *c, int n, const int *nodes /* [2][n] */,
                            const double *par /* [3][n] C, m, ic */);
/* BSIM3v3.3.0 instances in reference list order, same three-table form as BSIM4 with the lists of
 * csrc/bsim3_fields.h: nodes [6][ninst] d g s b d' s', flags [ninst] (B3F_*), prow, inst [NI][ninst],
 * mtab [nrows][NM], ptab [nrows][NP] -- replaces the BSIM3instance/BSIM3model walk of BSIM3load
 * (bsim3/b3ld.c:182-186).  nqsMod and acmMod != 0 return E_UNSUPP */
int ngbCircuitAddBsim3(ngb_circuit *c, int ninst, const int *nodes, const int *flags, const int *prow,
                       const double *inst, int nrows, const double *mtab, const double *ptab);
void ngbBsim3Layout(int out[6]);               /* model, bin, instance, node roles, stamp rows, states */
/* junction diodes after DIOsetup/DIOtemp (dio/diosetup.c, diotemp.c): nodes [6][n] pos neg posPrime
 * posSwPrime temp qp (posPrime == pos without series resistance; posSwPrime only with a separate sidewall
 * diode; temp != 0 exactly when flags has DIOF_SELFHEAT, qp != 0 exactly when it has DIOF_REVREC),
 * flags [n] DIOF_* and par [DIOP_COUNT][n] as listed in csrc/dio_fields.h (the raw model/instance rows
 * at the end are read only by instances with self-heating, whose load maps the parameters to
 * DIOtemp + delTemp every iteration like DIOtempUpdate, diotemp.c:18) -- replaces the
 * DIOinstance/DIOmodel walk of DIOload (dio/dioload.c:75-80) */
int ngbCircuitAddDiodes(ngb_circuit *c, int n, const int *nodes, const int *flags, const double *par);
/* VBIC bipolar transistors (4-terminal) after VBICsetup/VBICtemp (vbic/vbicsetup.c, vbictemp.c):
 * nodes [14][n] coll base emit subs collCX collCI baseBX baseBI emitEI baseBP subsSI temp xf1 xf2 (an
 * internal node equals its terminal when the series resistance is absent; temp != 0 exactly with the
 * self-heating flag, xf1/xf2 != 0 exactly with the excess-phase flag), flags [n] (0x1 off, 0x2 self-heating,
 * 0x4 excess phase), par [108][n] the parameter vector exactly as VBICload assembles it per instance
 * (vbic/vbicload.c:127-166), aux [6][n] type, tVcrit, icVBE, icVCE, area*m, temp -- replaces the
 * VBICinstance/VBICmodel walk of VBICload (vbicload.c:100-107).  ngbVbicLayout: [0]=parameters [1]=aux [2]=node roles [3]=states [4]=stamp rows */
int ngbCircuitAddVbic(ngb_circuit *c, int n, const int *nodes, const int *flags, const double *par, const double *aux);
void ngbVbicLayout(int out[5]);
int ngbCircuitAddVsources(ngb_circuit *c, int n, const int *nodes /* [3][n] pos neg branch */,
                          const int *fn /* [3][n] type order dcGiven */, const double *par /* [9][n] */);
/* PWL voltage source `inst` (of the ngbCircuitAddVsources table, type 5): the corner list VSRCcoeffs
 * (ncoef = VSRCfunctionOrder values t0 v0 t1 v1 ...), VSRCrdelay, and VSRCrBreakpt when `r=` makes the list
 * repeat (-1 otherwise) -- vsrc/vsrcload.c:324-367, breakpoints vsrcacct.c:174-226 */
int ngbCircuitSetVsourcePwl(ngb_circuit *c, int inst, int ncoef, const double *coef, double rdelay, int rbreakpt);
int ngbCircuitAddIsources(ngb_circuit *c, int n, const int *nodes /* [2][n] */,
                          const int *fn /* [3][n] */, const double *par /* [10][n] */);
/* PWL current source `inst` (of the ngbCircuitAddIsources table, type 5): the corner list ISRCcoeffs (ncoef =
 * ISRCfunctionOrder values t0 i0 t1 i1 ...).  ISRCload's PWL knows no delay and no repetition (isrc/isrcload.c:291-314)
 * and ISRCaccept sets the next corner only on a breakpoint (isrcacct.c:181-197); SFFM of a current source reads its
 * phases from coefficients 5 and 6 and applies no delay (isrcload.c:206-254).  All mirrored as they are */
int ngbCircuitSetIsourcePwl(ngb_circuit *c, int inst, int ncoef, const double *coef);

/* SMPmakeElt + SMPconvertCOOtoCSC + DEVbindCSC (klusmp.c:137-323, 417-440): builds the CSC
 * pattern, the slot map and the per-target contribution lists */
int ngbCircuitFinalize(ngb_circuit *c);
int ngbCircuitPatternSize(const ngb_circuit *c, int *n, int *nnz, int *nstamp_rows);
int ngbCircuitGetPatternEquations(const ngb_circuit *c, int *eq /* [n] equation number of pattern index k */);
int ngbCircuitGetPattern(const ngb_circuit *c, int *Ap, int *Ai, int *diag_slot /* [n] */);
int ngbCircuitGetBsim4Slots(const ngb_circuit *c, int *slots /* [70][ninst], -1 = ground/absent */);

/* SMPpreOrder + first SMPreorder: reuse the KLU symbolic analysis and pivot order
 * (klu_analyze, klu_factor) -- arrays exactly as in klu_symbolic / klu_numeric, with L/U column
 * patterns flattened (Lp/Li, Up/Ui in pivotal numbering) */
int ngbCircuitSetLuPattern(ngb_circuit *c, int n, int nblocks, const int *Q, const int *R,
                           const int *Pnum, const int *Lp, const int *Li, const int *Up, const int *Ui,
                           const int *Offp, const int *Offi);
/* NIiter runs the pivoting factor twice in a run with a DC operating point: at MODEINITJCT and in the
 * first iteration under MODEINITTRAN (niiter.c:107-111), and the pivot order may change.  Set 0
 * (default) holds the first result, set 1 the second; ngbCircuitSelectLuSet chooses which one the
 * next ngbCircuitSetLuPattern / ngbCircuitLuInfo refers to.  With only set 0 present it serves both */
/* .nodeset (kind 0) / .ic (kind 1) rows: eq [n] equation numbers, value [n] volts -- what CKTic (cktic.c) leaves in
 * CKTnode.nsGiven/nodeset, icGiven/ic.  Applied at the end of every load while the operating point is computed,
 * exactly as cktload.c:118-172 (ZeroNoncurRow :182-201).  Per-sample values: array "node.override" [n][S].
 * Call after ngbCircuitFinalize and before ngbBatchCreate */
int ngbCircuitSetNodeOverrides(ngb_circuit *c, int n, const int *eq, const int *kind, const double *value);
int ngbCircuitSelectLuSet(ngb_circuit *c, int which);
/* NIiter re-pivots (SMPreorder) at up to four moments of a run: [0] the MODEINITJCT iteration, [1] the
 * iteration after it (NISHOULDREORDER, niiter.c:335) whose order serves the rest of the operating point,
 * [2] the first iteration under MODEINITTRAN (niiter.c:107-111), [3] the iteration after it (:343-344),
 * whose order serves the rest of the transient.  set_of_event[4] names the pattern set (0..3, filled
 * through ngbCircuitSelectLuSet + ngbCircuitSetLuPattern) each of these factors produced; a UIC run
 * starts at [2].  Default without this call: set 0 for [0],[1]; set 1 (if filled, else 0) for [2],[3] */
int ngbCircuitSetLuEvents(ngb_circuit *c, const int *set_of_event);
/* ---- own pivoting factor (the role klu_factor plays behind SMPreorder, klusmp.c:700-760, klu_factor.c:384,
 * klu_kernel.c:642; csrc/ngb_pivot.c) ----
 * ngbCircuitSetSymbolic imports what klu_analyze leaves (klu_symbolic P, Q, R: block triangular form with a
 * fill-reducing order per block); ngbCircuitAnalyze computes an own one instead (one block, greedy minimum degree on
 * A + A': results then agree with the reference to rounding, not bit for bit).  ngbCircuitFactor runs the pivoting
 * left-looking factorization (threshold partial pivoting with diagonal preference, `pivtol` = CKTpivotRelTol, <= 0: 1e-3)
 * on one sample's matrix values and makes the result the pattern set selected by ngbCircuitSelectLuSet, exactly as if
 * it had been passed to ngbCircuitSetLuPattern; E_SINGULAR when a pivot is zero.  Inside ngbTranRun the same routine
 * re-pivots a sample whose refactor met a zero pivot (niiter.c:162-195). */
int ngbCircuitSetSymbolic(ngb_circuit *c, int n, int nblocks, const int *P, const int *Q, const int *R);
int ngbCircuitAnalyze(ngb_circuit *c);
int ngbCircuitFactor(ngb_circuit *c, const double *Ax /* [nnz], CSC slot order */, double pivtol);
int ngbCircuitGetLuPattern(const ngb_circuit *c, int *Pnum, int *Lp, int *Li, int *Up, int *Ui, int *Offp, int *Offi);
/* info: nV nlev npairs ntask nslev nsolvepairs lnz unz nzoff */
int ngbCircuitLuInfo(const ngb_circuit *c, int info[9]);

/* ---- batch ---- */
ngb_batch *ngbBatchCreate(ngb_circuit *c, int nsamples, int device);
void ngbBatchDestroy(ngb_batch *b);
/* after ngbCircuitSetLuPattern on a circuit that already has batches (a later SMPreorder): upload the
 * new pattern sets */
int ngbBatchRefreshLu(ngb_batch *b);

/* named device arrays (tests, the reference-side shim, result download):
 *   ctl.mode ctl.active ctl.head ctl.order ctl.noncon ctl.xsel ctl.err ctl.lusel int  [S]
 *   ctl.ag0 ctl.ag1 ctl.delta ctl.time ctl.gmin ctl.diag_gmin ctl.srcfact        f64   [S]
 *   ctl.delta_old                                                                f64   [7][S]
 *   x            f64 [2][neq+1][S]      Ax  f64 [S][nnz]      stamp f64 [rows][S]
 *   b4.inst      f64 [NI][ninst*S]      b4.state f64 [4][29][ninst*S]   b4.op f64 [NO][ninst*S]
 *   b4.prow      int [ninst*S]          cap.state f64 [4][2][ncap*S]    cap.par f64 [3][ncap*S]
 *   b3.inst      f64 [NI][n3*S]         b3.state f64 [4][17][n3*S]      b3.von f64 [n3*S]
 *   dio.par      f64 [NP][nd*S]         dio.state f64 [4][22][nd*S]
 *   vbic.par     f64 [108][nq*S]        vbic.aux f64 [6][nq*S]          vbic.state f64 [4][86][nq*S]
 *   vsrc.par     f64 [9][nv*S]          lu.V f64 [S][nV]    lu.Rs f64 [S][n]
 *   lu.nodeconv  int [S]                lu.singular int [S] */
long ngbBatchArrayBytes(ngb_batch *b, const char *name);
int ngbBatchUpload(ngb_batch *b, const char *name, const void *host, long bytes, long offset);
int ngbBatchDownload(ngb_batch *b, const char *name, void *host, long bytes, long offset);
void *ngbBatchDevPtr(ngb_batch *b, const char *name);
/* per-sample resistor values for parameter sweeps: g [nres][S] = RESconduct as REStemp computes it
 * (res/restemp.c), replaces `alter r = value` between runs */
int ngbBatchSetResistors(ngb_batch *b, const double *g);
int ngbBatchSetOpFull(ngb_batch *b, int on);   /* export every B4O_* field (parity runs) */

/* hot path, one call = one step of NIiter for every active sample */
int ngbLoad(ngb_batch *b);                     /* CKTload: device loads + assembly of Ax / rhs   */
int ngbLuFac(ngb_batch *b);                    /* SMPluFac: LoadGmin + row scaling + refactor     */
int ngbSolve(ngb_batch *b);                    /* SMPsolve + node part of NIconvTest              */
int ngbLuFacSolve(ngb_batch *b);               /* the two above in one launch                     */
int ngbNewtonStep(ngb_batch *b);               /* ngbLoad + ngbLuFacSolve                         */

/* device-resident transient analysis for the whole batch (DCtran + NIiter, per sample) */
int ngbTranRun(ngb_batch *b, int max_points, const int *save_eq, int nsave);
int ngbTranStats(ngb_batch *b, int *accepted, int *rejected, int *numiter, int *npoints /* each [S] */);
/* DCtran's return value per sample (dctran.c: 0, E_ITERLIM 103 when CKTop and its fallbacks fail, E_TIMESTEP 106 "timestep
 * too small", E_SINGULAR 102 ...): a failed sample stops, the others run on, ngbTranRun returns 0 */
int ngbTranErrors(ngb_batch *b, int *err /* [S] */);
long ngbTranWaveBytes(ngb_batch *b);
int ngbTranWaves(ngb_batch *b, double *times /* [S][max_points] */, double *values /* [S][max_points][nsave] */);
/* batch-aware binary rawfile of the stored waveforms (replaces the per-point OUTpData -> fileAddRealValue path of
 * src/frontend/outitf.c:633-765 and its header writers fileInit :881-923 / fileInit_pass2 :997-1029 for a batch): samples
 * first_sample .. first_sample + nsamples - 1 as consecutive `Transient Analysis` plots of one file, each with the reference's
 * header lines and `Binary:` rows {time, saved equations}.  names / types [nsave]: e.g. "v(out)" / "voltage"; date NULL = now */
int ngbTranWriteRaw(ngb_batch *b, const char *path, const char *title, const char *date, const char *const *names,
                    const char *const *types, int first_sample, int nsamples);
/* `.meas tran` on the device (the output path of src/frontend/outitf.c:633 + com_measure2.c:378-663 for this workload):
 * clause k is the count[k]-th RISE (kind 0) / FALL (1) / CROSS (2) of equation eq[k] through val[k], linearly
 * interpolated between the two accepted points around it, points before td[k] ignored -- evaluated as the points are
 * produced, so no waveform has to be stored or copied (max_points = 0, nsave = 0 is allowed then).  A
 * `TRIG .. TARG ..` measurement is two clauses; its result is the difference.  ngbTranMeasures: out [n][S], NaN = not found */
int ngbTranSetMeasures(ngb_batch *b, int n, const int *eq, const int *kind, const int *count, const double *val, const double *td);
int ngbTranMeasures(ngb_batch *b, double *out);
/* BSIM4temp inside the library (csrc/ngb_b4temp.c; b4temp.c:69-2408 + the clamps of b4check.c + b4geo.c): the load's model /
 * bin / instance tables from model cards and instance geometry, for any parameter value (continuous model-parameter
 * mismatch).  Tables are indexed by the name lists of csrc/bsim4_temp_fields.h: ngbBsim4TempLayout gives their lengths
 * {model, size, instance, binned}, ngbBsim4TempFieldName(list 0 model / 1 size / 2 instance, i) the names.
 *   temp: circuit temperature (K); vt0: CONSTvt0
 *   model [nmodel][NM] in/out (cards after BSIM4setup, `...Given` flags 0 / 1), inst [ninst][NI] in/out, inst_model [ninst]
 *   out: prow [ninst], *nrows, mtab [<= ninst rows][78], ptab [<= ninst rows][143], itab [51][ninst] -- what ngbCircuitAddBsim4
 *   and ngbBatchSetBsim4Rows take.  Returns 0, or NGB_E_PANIC where the reference stops with a fatal parameter error */
void ngbBsim4TempLayout(int layout[4]);
const char *ngbBsim4TempFieldName(int list, int i);
int ngbBsim4Temp(double temp, double vt0, int nmodel, double *model, int ninst, const int *inst_model, double *inst,
                 int *prow, int *nrows, double *mtab, double *ptab, double *itab);
/* direct ngbLoad calls: also evaluate DEVtrunc's step bounds into ctl.lte / ctl.lte2 (off by default -- the caller's own CKTtrunc
 * works on the host state vectors; ngbTranRun has its own arrangement, BSIM4trunc in a launch after the solve) */
void ngbBatchSetLoadLte(ngb_batch *b, int on);
/* which BSIM4 load kernel the batch runs (csrc/bsim4_variants.h): key[0] = the variant key packed from the model selectors
 * and rbodyMod / rgateMod of its instances (0xffffffff when they differ), key[1] = 1 when the kernel specialised on that key
 * is in use.  ngbBatchSetBsim4Generic(b, 1) (or NGB_B4_GENERIC=1 in the environment) forces the generic kernel: same bits */
/* per-sample parameter rows (ngbBatchSetBsim4Rows) are read as an overlay: only the columns that differ between the samples
 * of an instance at the thread's own row, the rest at the warp-uniform row of sample 0.  Returns 1 and the number of such
 * model / bin columns, or 0 (counts -1) without per-sample rows or with NGB_B4_OVERLAY=0 */
int ngbBatchBsim4Overlay(ngb_batch *b, int *model_columns, int *bin_columns);
int ngbBatchBsim4Variant(ngb_batch *b, unsigned key[2]);
void ngbBatchSetBsim4Generic(ngb_batch *b, int on);
/* per-stage device time of the Newton steps sampled by ngbProfile (ms summed over the sampled steps; returns their number):
 * [1] small device loads, [2] BSIM4 load, [3] assembly, [4] refactor + solve, [5] BSIM4trunc, [6] controller */
int ngbProfileStages(double ms[8]);
long ngbTranTicks(ngb_batch *b);              /* Newton steps the batch needed */
/* samples x events the host had to factor with pivoting in the last run (a zero pivot, or a pivoting event whose recorded
 * order failed the device's check of KLU's pivot rule on the sample's own matrix) */
int ngbTranRepivots(ngb_batch *b);
void *ngbTranDevWaves(ngb_batch *b, int which /* 0 times, 1 values */);   /* device pointers for a collective gather */
/* per-thread BSIM4 parameter rows for model-parameter mismatch: prow_t [ninst*S] into new tables */
int ngbBatchSetBsim4Rows(ngb_batch *b, const int *prow_t, int nrows, const double *mtab, const double *ptab);

#ifdef __cplusplus
}
#endif
#endif
