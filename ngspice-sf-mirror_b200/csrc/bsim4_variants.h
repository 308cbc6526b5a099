/* bsim4_variants.h -- compile-time specialisation of the BSIM4 load on the model selectors.
 *
 * The evaluation of b4ld.c branches on a dozen integer selectors that are constants of a model card (mobMod, capMod,
 * rdsMod, igcMod, ... -- b4set.c:95-330) or of an instance (rbodyMod, rgateMod) and on the integration method.  One
 * thread walks one instance, the load is bound by code volume and register pressure (profiles/README.md), so a kernel
 * that contains only the paths one card can take is smaller (14.3 k instead of 19.3 k SASS instructions for the
 * ro_17_4.cir cards) and spills half as much.  The selectors of a circuit are packed into a VARIANT KEY; the library
 * carries the generic kernel (key NGB_B4_GENERIC: every selector read from the parameter row, as before) and one
 * specialised instantiation per key listed in NGB_B4_VARIANT_KEYS.  A batch whose instances all share a listed key
 * runs the specialised kernel, every other batch the generic one; both execute the same floating-point operations
 * in the same order (the selected branches are the same source lines), so the results are the same bits.
 *
 * Key layout (bit position, width):                                                                            */
#ifndef NGB_BSIM4_VARIANTS_H
#define NGB_BSIM4_VARIANTS_H

#define B4K_mobMod_SH         0   /* 3 bits: 0..6 */
#define B4K_mobMod_W          3
#define B4K_capMod_SH         3   /* 2 */
#define B4K_capMod_W          2
#define B4K_cvchargeMod_SH    5   /* 1 */
#define B4K_cvchargeMod_W     1
#define B4K_dioMod_SH         6   /* 2 */
#define B4K_dioMod_W          2
#define B4K_rdsMod_SH         8   /* 1 */
#define B4K_rdsMod_W          1
#define B4K_igcMod_SH         9   /* 2 */
#define B4K_igcMod_W          2
#define B4K_igbMod_SH        11   /* 1 */
#define B4K_igbMod_W          1
#define B4K_gidlMod_SH       12   /* 1 */
#define B4K_gidlMod_W         1
#define B4K_tempMod_SH       13   /* 2 */
#define B4K_tempMod_W         2
#define B4K_mtrlMod_SH       15   /* 1 */
#define B4K_mtrlMod_W         1
#define B4K_mtrlCompatMod_SH 16   /* 1 */
#define B4K_mtrlCompatMod_W   1
#define B4K_rbodyMod_SH      17   /* 2 */
#define B4K_rbodyMod_W        2
#define B4K_rgateMod_SH      19   /* 2 */
#define B4K_rgateMod_W        2
#define B4K_gear_SH          21   /* 1: CKTintegrateMethod == GEAR */
#define B4K_gear_W            1
#define B4K_rowsO_SH         22   /* 1: per-sample parameter rows read as an overlay (bsim4_eval.cuh, B4OVL) */
#define B4K_rowsO_W           1

#define NGB_B4_GENERIC 0xffffffffu

#define B4K_FIELD(key, f) ((int)(((key) >> B4K_##f##_SH) & ((1u << B4K_##f##_W) - 1u)))
#define B4K_PACK(f, v)    (((unsigned)(v) & ((1u << B4K_##f##_W) - 1u)) << B4K_##f##_SH)
#define B4K_FITS(f, v)    ((v) >= 0 && (unsigned)(v) < (1u << B4K_##f##_W))

#define NGB_B4_KEY(v_mob, v_cap, v_cvchg, v_dio, v_rds, v_igc, v_igb, v_gidl, v_temp, v_mtrl, v_mtrlc, v_rbody, v_rgate, v_gear) \
    (B4K_PACK(mobMod, v_mob) | B4K_PACK(capMod, v_cap) | B4K_PACK(cvchargeMod, v_cvchg) | B4K_PACK(dioMod, v_dio) | \
     B4K_PACK(rdsMod, v_rds) | B4K_PACK(igcMod, v_igc) | B4K_PACK(igbMod, v_igb) | B4K_PACK(gidlMod, v_gidl) | \
     B4K_PACK(tempMod, v_temp) | B4K_PACK(mtrlMod, v_mtrl) | B4K_PACK(mtrlCompatMod, v_mtrlc) | \
     B4K_PACK(rbodyMod, v_rbody) | B4K_PACK(rgateMod, v_rgate) | B4K_PACK(gear, v_gear))

/* the specialised instantiations built into the library: X(key).  The cards the reference ships with its examples:
 *   examples/mos/ro_17_4.cir (N1 / P1): mobMod 0 capMod 2 dioMod 1 rdsMod 0 igcMod 1 igbMod 1 rbodyMod 1 rgateMod 1
 *   (the QA cards of tests/bsim4/{nmos,pmos}/parameters resolve to the same key)
 * more keys: add a line here (each costs ~15 s of compile time and ~230 KB of code) */
#define NGB_B4_VARIANT_KEYS(X) \
    X(NGB_B4_KEY(0, 2, 0, 1, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0)) \
    X(NGB_B4_KEY(0, 2, 0, 1, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0) | B4K_PACK(rowsO, 1))

/* the variant a batch runs: its key when the library carries that instantiation, NGB_B4_GENERIC otherwise */
static inline int b4_variant_built(unsigned key)
{
#define X(k) if (key == (k)) return 1;
    NGB_B4_VARIANT_KEYS(X)
#undef X
    return 0;
}

#endif
