/* ngb_math.cuh -- exp() and log() that round exactly like the libm the reference runs on.
 *
 * ngspice calls libm's exp/log ~40 times per BSIM4 evaluation, and a ring oscillator started
 * from its metastable point amplifies a last-place difference in any of them into a different
 * waveform.  libm is the one arithmetic dependency outside the reference tree (SURVEY.md 8c:
 * "its last-ulp behaviour is part of the oracle's results"), so the device implements the SAME
 * published algorithm with the SAME tables: the table-driven exp/log of ARM's "Optimized
 * Routines" as shipped in glibc >= 2.28 (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c; here glibc
 * 2.39, x86-64 FMA variant -- the fused operations below are the ones that variant executes).
 * Checked bit-for-bit against that libm on 3e7 random arguments (tests/test_math_replica.py);
 * every a*b+c written as fma() is a single rounding on both the CPU and the GPU.
 *
 * Arguments outside the main path (|x| >= 512 or < 2^-54 for exp; x <= 0, inf, nan or subnormal
 * for log) take the platform's own exp/log -- the model code clamps its arguments long before.
 */
#ifndef NGB_MATH_CUH
#define NGB_MATH_CUH
#include "ngb_common.h"
#include <string.h>

#ifdef __CUDACC__
#define NGB_TABLE __device__ const
#else
#define NGB_TABLE static const
#endif

/* On the device exp/log are real (out-of-line) functions: the BSIM4 evaluation is bound by
 * instruction fetch (18k straight-line instructions per evaluation, far beyond the instruction
 * caches), and ~70 inlined copies of these bodies would be streamed from L2 by every warp; one
 * shared copy stays cache-resident. */
#ifdef __CUDACC__
#define NGB_MATH_FN __device__ __noinline__
#else
#define NGB_MATH_FN static inline
#endif

/* 2^(i/128) as (tail, scale-bits) pairs */
NGB_TABLE unsigned long long ngb_exp_tab[256] = {
    0x0000000000000000ULL, 0x3ff0000000000000ULL, 0x3c9b3b4f1a88bf6eULL, 0x3feff63da9fb3335ULL,
    0xbc7160139cd8dc5dULL, 0x3fefec9a3e778061ULL, 0xbc905e7a108766d1ULL, 0x3fefe315e86e7f85ULL,
    0x3c8cd2523567f613ULL, 0x3fefd9b0d3158574ULL, 0xbc8bce8023f98efaULL, 0x3fefd06b29ddf6deULL,
    0x3c60f74e61e6c861ULL, 0x3fefc74518759bc8ULL, 0x3c90a3e45b33d399ULL, 0x3fefbe3ecac6f383ULL,
    0x3c979aa65d837b6dULL, 0x3fefb5586cf9890fULL, 0x3c8eb51a92fdeffcULL, 0x3fefac922b7247f7ULL,
    0x3c3ebe3d702f9cd1ULL, 0x3fefa3ec32d3d1a2ULL, 0xbc6a033489906e0bULL, 0x3fef9b66affed31bULL,
    0xbc9556522a2fbd0eULL, 0x3fef9301d0125b51ULL, 0xbc5080ef8c4eea55ULL, 0x3fef8abdc06c31ccULL,
    0xbc91c923b9d5f416ULL, 0x3fef829aaea92de0ULL, 0x3c80d3e3e95c55afULL, 0x3fef7a98c8a58e51ULL,
    0xbc801b15eaa59348ULL, 0x3fef72b83c7d517bULL, 0xbc8f1ff055de323dULL, 0x3fef6af9388c8deaULL,
    0x3c8b898c3f1353bfULL, 0x3fef635beb6fcb75ULL, 0xbc96d99c7611eb26ULL, 0x3fef5be084045cd4ULL,
    0x3c9aecf73e3a2f60ULL, 0x3fef54873168b9aaULL, 0xbc8fe782cb86389dULL, 0x3fef4d5022fcd91dULL,
    0x3c8a6f4144a6c38dULL, 0x3fef463b88628cd6ULL, 0x3c807a05b0e4047dULL, 0x3fef3f49917ddc96ULL,
    0x3c968efde3a8a894ULL, 0x3fef387a6e756238ULL, 0x3c875e18f274487dULL, 0x3fef31ce4fb2a63fULL,
    0x3c80472b981fe7f2ULL, 0x3fef2b4565e27cddULL, 0xbc96b87b3f71085eULL, 0x3fef24dfe1f56381ULL,
    0x3c82f7e16d09ab31ULL, 0x3fef1e9df51fdee1ULL, 0xbc3d219b1a6fbffaULL, 0x3fef187fd0dad990ULL,
    0x3c8b3782720c0ab4ULL, 0x3fef1285a6e4030bULL, 0x3c6e149289cecb8fULL, 0x3fef0cafa93e2f56ULL,
    0x3c834d754db0abb6ULL, 0x3fef06fe0a31b715ULL, 0x3c864201e2ac744cULL, 0x3fef0170fc4cd831ULL,
    0x3c8fdd395dd3f84aULL, 0x3feefc08b26416ffULL, 0xbc86a3803b8e5b04ULL, 0x3feef6c55f929ff1ULL,
    0xbc924aedcc4b5068ULL, 0x3feef1a7373aa9cbULL, 0xbc9907f81b512d8eULL, 0x3feeecae6d05d866ULL,
    0xbc71d1e83e9436d2ULL, 0x3feee7db34e59ff7ULL, 0xbc991919b3ce1b15ULL, 0x3feee32dc313a8e5ULL,
    0x3c859f48a72a4c6dULL, 0x3feedea64c123422ULL, 0xbc9312607a28698aULL, 0x3feeda4504ac801cULL,
    0xbc58a78f4817895bULL, 0x3feed60a21f72e2aULL, 0xbc7c2c9b67499a1bULL, 0x3feed1f5d950a897ULL,
    0x3c4363ed60c2ac11ULL, 0x3feece086061892dULL, 0x3c9666093b0664efULL, 0x3feeca41ed1d0057ULL,
    0x3c6ecce1daa10379ULL, 0x3feec6a2b5c13cd0ULL, 0x3c93ff8e3f0f1230ULL, 0x3feec32af0d7d3deULL,
    0x3c7690cebb7aafb0ULL, 0x3feebfdad5362a27ULL, 0x3c931dbdeb54e077ULL, 0x3feebcb299fddd0dULL,
    0xbc8f94340071a38eULL, 0x3feeb9b2769d2ca7ULL, 0xbc87deccdc93a349ULL, 0x3feeb6daa2cf6642ULL,
    0xbc78dec6bd0f385fULL, 0x3feeb42b569d4f82ULL, 0xbc861246ec7b5cf6ULL, 0x3feeb1a4ca5d920fULL,
    0x3c93350518fdd78eULL, 0x3feeaf4736b527daULL, 0x3c7b98b72f8a9b05ULL, 0x3feead12d497c7fdULL,
    0x3c9063e1e21c5409ULL, 0x3feeab07dd485429ULL, 0x3c34c7855019c6eaULL, 0x3feea9268a5946b7ULL,
    0x3c9432e62b64c035ULL, 0x3feea76f15ad2148ULL, 0xbc8ce44a6199769fULL, 0x3feea5e1b976dc09ULL,
    0xbc8c33c53bef4da8ULL, 0x3feea47eb03a5585ULL, 0xbc845378892be9aeULL, 0x3feea34634ccc320ULL,
    0xbc93cedd78565858ULL, 0x3feea23882552225ULL, 0x3c5710aa807e1964ULL, 0x3feea155d44ca973ULL,
    0xbc93b3efbf5e2228ULL, 0x3feea09e667f3bcdULL, 0xbc6a12ad8734b982ULL, 0x3feea012750bdabfULL,
    0xbc6367efb86da9eeULL, 0x3fee9fb23c651a2fULL, 0xbc80dc3d54e08851ULL, 0x3fee9f7df9519484ULL,
    0xbc781f647e5a3ecfULL, 0x3fee9f75e8ec5f74ULL, 0xbc86ee4ac08b7db0ULL, 0x3fee9f9a48a58174ULL,
    0xbc8619321e55e68aULL, 0x3fee9feb564267c9ULL, 0x3c909ccb5e09d4d3ULL, 0x3feea0694fde5d3fULL,
    0xbc7b32dcb94da51dULL, 0x3feea11473eb0187ULL, 0x3c94ecfd5467c06bULL, 0x3feea1ed0130c132ULL,
    0x3c65ebe1abd66c55ULL, 0x3feea2f336cf4e62ULL, 0xbc88a1c52fb3cf42ULL, 0x3feea427543e1a12ULL,
    0xbc9369b6f13b3734ULL, 0x3feea589994cce13ULL, 0xbc805e843a19ff1eULL, 0x3feea71a4623c7adULL,
    0xbc94d450d872576eULL, 0x3feea8d99b4492edULL, 0x3c90ad675b0e8a00ULL, 0x3feeaac7d98a6699ULL,
    0x3c8db72fc1f0eab4ULL, 0x3feeace5422aa0dbULL, 0xbc65b6609cc5e7ffULL, 0x3feeaf3216b5448cULL,
    0x3c7bf68359f35f44ULL, 0x3feeb1ae99157736ULL, 0xbc93091fa71e3d83ULL, 0x3feeb45b0b91ffc6ULL,
    0xbc5da9b88b6c1e29ULL, 0x3feeb737b0cdc5e5ULL, 0xbc6c23f97c90b959ULL, 0x3feeba44cbc8520fULL,
    0xbc92434322f4f9aaULL, 0x3feebd829fde4e50ULL, 0xbc85ca6cd7668e4bULL, 0x3feec0f170ca07baULL,
    0x3c71affc2b91ce27ULL, 0x3feec49182a3f090ULL, 0x3c6dd235e10a73bbULL, 0x3feec86319e32323ULL,
    0xbc87c50422622263ULL, 0x3feecc667b5de565ULL, 0x3c8b1c86e3e231d5ULL, 0x3feed09bec4a2d33ULL,
    0xbc91bbd1d3bcbb15ULL, 0x3feed503b23e255dULL, 0x3c90cc319cee31d2ULL, 0x3feed99e1330b358ULL,
    0x3c8469846e735ab3ULL, 0x3feede6b5579fdbfULL, 0xbc82dfcd978e9db4ULL, 0x3feee36bbfd3f37aULL,
    0x3c8c1a7792cb3387ULL, 0x3feee89f995ad3adULL, 0xbc907b8f4ad1d9faULL, 0x3feeee07298db666ULL,
    0xbc55c3d956dcaebaULL, 0x3feef3a2b84f15fbULL, 0xbc90a40e3da6f640ULL, 0x3feef9728de5593aULL,
    0xbc68d6f438ad9334ULL, 0x3feeff76f2fb5e47ULL, 0xbc91eee26b588a35ULL, 0x3fef05b030a1064aULL,
    0x3c74ffd70a5fddcdULL, 0x3fef0c1e904bc1d2ULL, 0xbc91bdfbfa9298acULL, 0x3fef12c25bd71e09ULL,
    0x3c736eae30af0cb3ULL, 0x3fef199bdd85529cULL, 0x3c8ee3325c9ffd94ULL, 0x3fef20ab5fffd07aULL,
    0x3c84e08fd10959acULL, 0x3fef27f12e57d14bULL, 0x3c63cdaf384e1a67ULL, 0x3fef2f6d9406e7b5ULL,
    0x3c676b2c6c921968ULL, 0x3fef3720dcef9069ULL, 0xbc808a1883ccb5d2ULL, 0x3fef3f0b555dc3faULL,
    0xbc8fad5d3ffffa6fULL, 0x3fef472d4a07897cULL, 0xbc900dae3875a949ULL, 0x3fef4f87080d89f2ULL,
    0x3c74a385a63d07a7ULL, 0x3fef5818dcfba487ULL, 0xbc82919e2040220fULL, 0x3fef60e316c98398ULL,
    0x3c8e5a50d5c192acULL, 0x3fef69e603db3285ULL, 0x3c843a59ac016b4bULL, 0x3fef7321f301b460ULL,
    0xbc82d52107b43e1fULL, 0x3fef7c97337b9b5fULL, 0xbc892ab93b470dc9ULL, 0x3fef864614f5a129ULL,
    0x3c74b604603a88d3ULL, 0x3fef902ee78b3ff6ULL, 0x3c83c5ec519d7271ULL, 0x3fef9a51fbc74c83ULL,
    0xbc8ff7128fd391f0ULL, 0x3fefa4afa2a490daULL, 0xbc8dae98e223747dULL, 0x3fefaf482d8e67f1ULL,
    0x3c8ec3bc41aa2008ULL, 0x3fefba1bee615a27ULL, 0x3c842b94c3a9eb32ULL, 0x3fefc52b376bba97ULL,
    0x3c8a64a931d185eeULL, 0x3fefd0765b6e4540ULL, 0xbc8e37bae43be3edULL, 0x3fefdbfdad9cbe14ULL,
    0x3c77893b4d91cd9dULL, 0x3fefe7c1819e90d8ULL, 0x3c5305c14160cc89ULL, 0x3feff3c22b8f71f1ULL,
};
/* log: (1/c, log(c)) for the 128 sub-intervals of [0.6875, 1.375) */
NGB_TABLE unsigned long long ngb_log_tab[256] = {
    0x3ff734f0c3e0de9fULL, 0xbfd7cc7f79e69000ULL, 0x3ff713786a2ce91fULL, 0xbfd76feec20d0000ULL,
    0x3ff6f26008fab5a0ULL, 0xbfd713e31351e000ULL, 0x3ff6d1a61f138c7dULL, 0xbfd6b85b38287800ULL,
    0x3ff6b1490bc5b4d1ULL, 0xbfd65d5590807800ULL, 0x3ff69147332f0cbaULL, 0xbfd602d076180000ULL,
    0x3ff6719f18224223ULL, 0xbfd5a8ca86909000ULL, 0x3ff6524f99a51ed9ULL, 0xbfd54f4356035000ULL,
    0x3ff63356aa8f24c4ULL, 0xbfd4f637c36b4000ULL, 0x3ff614b36b9ddc14ULL, 0xbfd49da7fda85000ULL,
    0x3ff5f66452c65c4cULL, 0xbfd445923989a800ULL, 0x3ff5d867b5912c4fULL, 0xbfd3edf439b0b800ULL,
    0x3ff5babccb5b90deULL, 0xbfd396ce448f7000ULL, 0x3ff59d61f2d91a78ULL, 0xbfd3401e17bda000ULL,
    0x3ff5805612465687ULL, 0xbfd2e9e2ef468000ULL, 0x3ff56397cee76bd3ULL, 0xbfd2941b3830e000ULL,
    0x3ff54725e2a77f93ULL, 0xbfd23ec58cda8800ULL, 0x3ff52aff42064583ULL, 0xbfd1e9e129279000ULL,
    0x3ff50f22dbb2bddfULL, 0xbfd1956d2b48f800ULL, 0x3ff4f38f4734ded7ULL, 0xbfd141679ab9f800ULL,
    0x3ff4d843cfde2840ULL, 0xbfd0edd094ef9800ULL, 0x3ff4bd3ec078a3c8ULL, 0xbfd09aa518db1000ULL,
    0x3ff4a27fc3e0258aULL, 0xbfd047e65263b800ULL, 0x3ff4880524d48434ULL, 0xbfcfeb224586f000ULL,
    0x3ff46dce1b192d0bULL, 0xbfcf474a7517b000ULL, 0x3ff453d9d3391854ULL, 0xbfcea4443d103000ULL,
    0x3ff43a2744b4845aULL, 0xbfce020d44e9b000ULL, 0x3ff420b54115f8fbULL, 0xbfcd60a22977f000ULL,
    0x3ff40782da3ef4b1ULL, 0xbfccc00104959000ULL, 0x3ff3ee8f5d57fe8fULL, 0xbfcc202956891000ULL,
    0x3ff3d5d9a00b4ce9ULL, 0xbfcb81178d811000ULL, 0x3ff3bd60c010c12bULL, 0xbfcae2c9ccd3d000ULL,
    0x3ff3a5242b75dab8ULL, 0xbfca45402e129000ULL, 0x3ff38d22cd9fd002ULL, 0xbfc9a877681df000ULL,
    0x3ff3755bc5847a1cULL, 0xbfc90c6d69483000ULL, 0x3ff35dce49ad36e2ULL, 0xbfc87120a645c000ULL,
    0x3ff34679984dd440ULL, 0xbfc7d68fb4143000ULL, 0x3ff32f5cceffcb24ULL, 0xbfc73cb83c627000ULL,
    0x3ff3187775a10d49ULL, 0xbfc6a39a9b376000ULL, 0x3ff301c8373e3990ULL, 0xbfc60b3154b7a000ULL,
    0x3ff2eb4ebb95f841ULL, 0xbfc5737d76243000ULL, 0x3ff2d50a0219a9d1ULL, 0xbfc4dc7b8fc23000ULL,
    0x3ff2bef9a8b7fd2aULL, 0xbfc4462c51d20000ULL, 0x3ff2a91c7a0c1babULL, 0xbfc3b08abc830000ULL,
    0x3ff293726014b530ULL, 0xbfc31b996b490000ULL, 0x3ff27dfa5757a1f5ULL, 0xbfc2875490a44000ULL,
    0x3ff268b39b1d3bbfULL, 0xbfc1f3b9f879a000ULL, 0x3ff2539d838ff5bdULL, 0xbfc160c8252ca000ULL,
    0x3ff23eb7aac9083bULL, 0xbfc0ce7f57f72000ULL, 0x3ff22a012ba940b6ULL, 0xbfc03cdc49fea000ULL,
    0x3ff2157996cc4132ULL, 0xbfbf57bdbc4b8000ULL, 0x3ff201201dd2fc9bULL, 0xbfbe370896404000ULL,
    0x3ff1ecf4494d480bULL, 0xbfbd17983ef94000ULL, 0x3ff1d8f5528f6569ULL, 0xbfbbf9674ed8a000ULL,
    0x3ff1c52311577e7cULL, 0xbfbadc79202f6000ULL, 0x3ff1b17c74cb26e9ULL, 0xbfb9c0c3e7288000ULL,
    0x3ff19e010c2c1ab6ULL, 0xbfb8a646b372c000ULL, 0x3ff18ab07bb670bdULL, 0xbfb78d01b3ac0000ULL,
    0x3ff1778a25efbcb6ULL, 0xbfb674f145380000ULL, 0x3ff1648d354c31daULL, 0xbfb55e0e6d878000ULL,
    0x3ff151b990275fddULL, 0xbfb4485cdea1e000ULL, 0x3ff13f0ea432d24cULL, 0xbfb333d94d6aa000ULL,
    0x3ff12c8b7210f9daULL, 0xbfb22079f8c56000ULL, 0x3ff11a3028ecb531ULL, 0xbfb10e4698622000ULL,
    0x3ff107fbda8434afULL, 0xbfaffa6c6ad20000ULL, 0x3ff0f5ee0f4e6bb3ULL, 0xbfadda8d4a774000ULL,
    0x3ff0e4065d2a9fceULL, 0xbfabbcece4850000ULL, 0x3ff0d244632ca521ULL, 0xbfa9a1894012c000ULL,
    0x3ff0c0a77ce2981aULL, 0xbfa788583302c000ULL, 0x3ff0af2f83c636d1ULL, 0xbfa5715e67d68000ULL,
    0x3ff09ddb98a01339ULL, 0xbfa35c8a49658000ULL, 0x3ff08cabaf52e7dfULL, 0xbfa149e364154000ULL,
    0x3ff07b9f2f4e28fbULL, 0xbf9e72c082eb8000ULL, 0x3ff06ab58c358f19ULL, 0xbf9a55f152528000ULL,
    0x3ff059eea5ecf92cULL, 0xbf963d62cf818000ULL, 0x3ff04949cdd12c90ULL, 0xbf9228fb8caa0000ULL,
    0x3ff038c6c6f0ada9ULL, 0xbf8c317b20f90000ULL, 0x3ff02865137932a9ULL, 0xbf8419355daa0000ULL,
    0x3ff0182427ea7348ULL, 0xbf781203c2ec0000ULL, 0x3ff008040614b195ULL, 0xbf60040979240000ULL,
    0x3fefe01ff726fa1aULL, 0x3f6feff384900000ULL, 0x3fefa11cc261ea74ULL, 0x3f87dc41353d0000ULL,
    0x3fef6310b081992eULL, 0x3f93cea3c4c28000ULL, 0x3fef25f63ceeadcdULL, 0x3f9b9fc114890000ULL,
    0x3feee9c8039113e7ULL, 0x3fa1b0d8ce110000ULL, 0x3feeae8078cbb1abULL, 0x3fa58a5bd001c000ULL,
    0x3fee741aa29d0c9bULL, 0x3fa95c8340d88000ULL, 0x3fee3a91830a99b5ULL, 0x3fad276aef578000ULL,
    0x3fee01e009609a56ULL, 0x3fb07598e598c000ULL, 0x3fedca01e577bb98ULL, 0x3fb253f5e30d2000ULL,
    0x3fed92f20b7c9103ULL, 0x3fb42edd8b380000ULL, 0x3fed5cac66fb5cceULL, 0x3fb606598757c000ULL,
    0x3fed272caa5ede9dULL, 0x3fb7da76356a0000ULL, 0x3fecf26e3e6b2ccdULL, 0x3fb9ab434e1c6000ULL,
    0x3fecbe6da2a77902ULL, 0x3fbb78c7bb0d6000ULL, 0x3fec8b266d37086dULL, 0x3fbd431332e72000ULL,
    0x3fec5894bd5d5804ULL, 0x3fbf0a3171de6000ULL, 0x3fec26b533bb9f8cULL, 0x3fc067152b914000ULL,
    0x3febf583eeece73fULL, 0x3fc147858292b000ULL, 0x3febc4fd75db96c1ULL, 0x3fc2266ecdca3000ULL,
    0x3feb951e0c864a28ULL, 0x3fc303d7a6c55000ULL, 0x3feb65e2c5ef3e2cULL, 0x3fc3dfc33c331000ULL,
    0x3feb374867c9888bULL, 0x3fc4ba366b7a8000ULL, 0x3feb094b211d304aULL, 0x3fc5933928d1f000ULL,
    0x3feadbe885f2ef7eULL, 0x3fc66acd2418f000ULL, 0x3feaaf1d31603da2ULL, 0x3fc740f8ec669000ULL,
    0x3fea82e63fd358a7ULL, 0x3fc815c0f51af000ULL, 0x3fea5740ef09738bULL, 0x3fc8e92954f68000ULL,
    0x3fea2c2a90ab4b27ULL, 0x3fc9bb3602f84000ULL, 0x3fea01a01393f2d1ULL, 0x3fca8bed1c2c0000ULL,
    0x3fe9d79f24db3c1bULL, 0x3fcb5b515c01d000ULL, 0x3fe9ae2505c7b190ULL, 0x3fcc2967ccbcc000ULL,
    0x3fe9852ef297ce2fULL, 0x3fccf635d5486000ULL, 0x3fe95cbaeea44b75ULL, 0x3fcdc1bd3446c000ULL,
    0x3fe934c69de74838ULL, 0x3fce8c01b8cfe000ULL, 0x3fe90d4f2f6752e6ULL, 0x3fcf5509c0179000ULL,
    0x3fe8e6528effd79dULL, 0x3fd00e6c121fb800ULL, 0x3fe8bfce9fcc007cULL, 0x3fd071b80e93d000ULL,
    0x3fe899c0dabec30eULL, 0x3fd0d46b9e867000ULL, 0x3fe87427aa2317fbULL, 0x3fd13687334bd000ULL,
    0x3fe84f00acb39a08ULL, 0x3fd1980d67234800ULL, 0x3fe82a49e8653e55ULL, 0x3fd1f8ffe0cc8000ULL,
    0x3fe8060195f40260ULL, 0x3fd2595fd7636800ULL, 0x3fe7e22563e0a329ULL, 0x3fd2b9300914a800ULL,
    0x3fe7beb377dcb5adULL, 0x3fd3187210436000ULL, 0x3fe79baa679725c2ULL, 0x3fd377266dec1800ULL,
    0x3fe77907f2170657ULL, 0x3fd3d54ffbaf3000ULL, 0x3fe756cadbd6130cULL, 0x3fd432eee32fe000ULL,
};

NGB_HD double ngb_bits2d(unsigned long long u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double f; memcpy(&f, &u, 8); return f;
#endif
}
NGB_HD unsigned long long ngb_d2bits(double f)
{
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(f);
#else
    unsigned long long u; memcpy(&u, &f, 8); return u;
#endif
}
#ifdef __CUDA_ARCH__
#define NGB_TAB(t, i) __ldg(&(t)[i])
#else
#define NGB_TAB(t, i) ((t)[i])
#endif

NGB_MATH_FN double ngb_exp(double x)
{
    const double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8000000000000p+52;
    const double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    const unsigned abstop = (unsigned)(ngb_d2bits(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;      /* |x| < 2^-54 */
        return exp(x);                                           /* |x| >= 512, inf, nan */
    }
    double kd = fma(x, InvLn2N, Shift);
    const unsigned long long ki = ngb_d2bits(kd);
    kd -= Shift;
    const double r = fma(kd, NegLn2loN, fma(kd, NegLn2hiN, x));
    const unsigned idx = 2u * (unsigned)(ki & 127u);
    const unsigned long long top = ki << 45;
    const double tail = ngb_bits2d(NGB_TAB(ngb_exp_tab, idx));
    const unsigned long long sbits = NGB_TAB(ngb_exp_tab, idx + 1) + top;
    const double r2 = r * r;
    const double p1 = fma(r, C3, C2), p2 = fma(r, C5, C4);
    const double tmp = fma(r2 * r2, p2, fma(r2, p1, tail + r));
    const double scale = ngb_bits2d(sbits);
    return fma(scale, tmp, scale);
}

NGB_MATH_FN double ngb_log(double x)
{
    const double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    const double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3, A3 = 0x1.999b324f10111p-3, A4 = -0x1.55575e506c89fp-3;
    const double B0 = -0x1.0000000000000p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3, B3 = 0x1.999999995dd0cp-3, B4 = -0x1.55555556745a7p-3,
                 B5 = 0x1.24924a344de30p-3, B6 = -0x1.fffffa4423d65p-4, B7 = 0x1.c7184282ad6cap-4, B8 = -0x1.999eb43b068ffp-4, B9 = 0x1.78182f7afd085p-4, B10 = -0x1.5521375d145cdp-4;
    const unsigned long long ix = ngb_d2bits(x);
    const unsigned top = (unsigned)(ix >> 48);
    const unsigned long long LO = 0x3fee000000000000ULL, HI = 0x3ff1090000000000ULL;   /* 1 - 2^-4, 1 + 0x1.09p-4 */
    if (ix - LO < HI - LO) {
        if (ix == 0x3ff0000000000000ULL) return 0.0;
        const double r = x - 1.0, r2 = r * r, r3 = r * r2;
        const double a = fma(r2, B3, fma(r, B2, B1));
        const double b = fma(r2, B6, fma(r, B5, B4));
        const double c = fma(r3, B10, fma(r2, B9, fma(r, B8, B7)));
        const double i1 = fma(fma(c, r3, b), r3, a);
        const double t = fma(r, 0x1p27, r);
        const double rhi = fma(-0x1p27, r, t), rlo = r - rhi;
        const double sq = rhi * rhi;
        const double hi = fma(sq, B0, r);
        const double lo = fma(sq, B0, r - hi);
        const double lo2 = fma(B0 * rlo, rhi + r, lo);
        return hi + fma(i1, r3, lo2);
    }
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) return log(x);       /* x <= 0, inf, nan, subnormal */
    const unsigned long long tmp = ix - 0x3fe6000000000000ULL;
    const unsigned i = (unsigned)(tmp >> 45) & 127u;
    const int k = (int)((long long)tmp >> 52);
    const unsigned long long iz = ix - (tmp & 0xfff0000000000000ULL);
    const double invc = ngb_bits2d(NGB_TAB(ngb_log_tab, 2 * i)), logc = ngb_bits2d(NGB_TAB(ngb_log_tab, 2 * i + 1));
    const double z = ngb_bits2d(iz);
    const double r = fma(z, invc, -1.0), kd = (double)k;
    const double w = fma(kd, Ln2hi, logc), hi = w + r;
    const double lo = fma(kd, Ln2lo, (w - hi) + r);
    const double r2 = r * r;
    const double p = fma(r2, fma(r, A4, A3), fma(r, A2, A1));
    return fma(r * r2, p, fma(r2, A0, lo)) + hi;
}
/* pow(): log_inline with the 128-entry table of glibc's __pow_log_data (invc, logc, logctail per
 * entry), then exp with the low part of y*log(x) carried along (e_pow.c of glibc 2.39, FMA variant,
 * operation order as executed by __pow_fma).  Main path: x positive and normal, 2^-65 <= |y| < 2^63,
 * 2^-54 <= |y*log x| < 512; everything else takes the platform's pow(). */
NGB_TABLE unsigned long long ngb_pow_tab[384] = {
    0x3ff6a00000000000ULL, 0xbfd62c82f2b9c800ULL, 0x3cfab42428375680ULL, 0x3ff6800000000000ULL,
    0xbfd5d1bdbf580800ULL, 0xbd1ca508d8e0f720ULL, 0x3ff6600000000000ULL, 0xbfd5767717455800ULL,
    0xbd2362a4d5b6506dULL, 0x3ff6400000000000ULL, 0xbfd51aad872df800ULL, 0xbce684e49eb067d5ULL,
    0x3ff6200000000000ULL, 0xbfd4be5f95777800ULL, 0xbd041b6993293ee0ULL, 0x3ff6000000000000ULL,
    0xbfd4618bc21c6000ULL, 0x3d13d82f484c84ccULL, 0x3ff5e00000000000ULL, 0xbfd404308686a800ULL,
    0x3cdc42f3ed820b3aULL, 0x3ff5c00000000000ULL, 0xbfd3a64c55694800ULL, 0x3d20b1c686519460ULL,
    0x3ff5a00000000000ULL, 0xbfd347dd9a988000ULL, 0x3d25594dd4c58092ULL, 0x3ff5800000000000ULL,
    0xbfd2e8e2bae12000ULL, 0x3d267b1e99b72bd8ULL, 0x3ff5600000000000ULL, 0xbfd2895a13de8800ULL,
    0x3d15ca14b6cfb03fULL, 0x3ff5600000000000ULL, 0xbfd2895a13de8800ULL, 0x3d15ca14b6cfb03fULL,
    0x3ff5400000000000ULL, 0xbfd22941fbcf7800ULL, 0xbd165a242853da76ULL, 0x3ff5200000000000ULL,
    0xbfd1c898c1699800ULL, 0xbd1fafbc68e75404ULL, 0x3ff5000000000000ULL, 0xbfd1675cababa800ULL,
    0x3d1f1fc63382a8f0ULL, 0x3ff4e00000000000ULL, 0xbfd1058bf9ae4800ULL, 0xbd26a8c4fd055a66ULL,
    0x3ff4c00000000000ULL, 0xbfd0a324e2739000ULL, 0xbd0c6bee7ef4030eULL, 0x3ff4a00000000000ULL,
    0xbfd0402594b4d000ULL, 0xbcf036b89ef42d7fULL, 0x3ff4a00000000000ULL, 0xbfd0402594b4d000ULL,
    0xbcf036b89ef42d7fULL, 0x3ff4800000000000ULL, 0xbfcfb9186d5e4000ULL, 0x3d0d572aab993c87ULL,
    0x3ff4600000000000ULL, 0xbfcef0adcbdc6000ULL, 0x3d2b26b79c86af24ULL, 0x3ff4400000000000ULL,
    0xbfce27076e2af000ULL, 0xbd172f4f543fff10ULL, 0x3ff4200000000000ULL, 0xbfcd5c216b4fc000ULL,
    0x3d21ba91bbca681bULL, 0x3ff4000000000000ULL, 0xbfcc8ff7c79aa000ULL, 0x3d27794f689f8434ULL,
    0x3ff4000000000000ULL, 0xbfcc8ff7c79aa000ULL, 0x3d27794f689f8434ULL, 0x3ff3e00000000000ULL,
    0xbfcbc286742d9000ULL, 0x3d194eb0318bb78fULL, 0x3ff3c00000000000ULL, 0xbfcaf3c94e80c000ULL,
    0x3cba4e633fcd9066ULL, 0x3ff3a00000000000ULL, 0xbfca23bc1fe2b000ULL, 0xbd258c64dc46c1eaULL,
    0x3ff3a00000000000ULL, 0xbfca23bc1fe2b000ULL, 0xbd258c64dc46c1eaULL, 0x3ff3800000000000ULL,
    0xbfc9525a9cf45000ULL, 0xbd2ad1d904c1d4e3ULL, 0x3ff3600000000000ULL, 0xbfc87fa06520d000ULL,
    0x3d2bbdbf7fdbfa09ULL, 0x3ff3400000000000ULL, 0xbfc7ab890210e000ULL, 0x3d2bdb9072534a58ULL,
    0x3ff3400000000000ULL, 0xbfc7ab890210e000ULL, 0x3d2bdb9072534a58ULL, 0x3ff3200000000000ULL,
    0xbfc6d60fe719d000ULL, 0xbd10e46aa3b2e266ULL, 0x3ff3000000000000ULL, 0xbfc5ff3070a79000ULL,
    0xbd1e9e439f105039ULL, 0x3ff3000000000000ULL, 0xbfc5ff3070a79000ULL, 0xbd1e9e439f105039ULL,
    0x3ff2e00000000000ULL, 0xbfc526e5e3a1b000ULL, 0xbd20de8b90075b8fULL, 0x3ff2c00000000000ULL,
    0xbfc44d2b6ccb8000ULL, 0x3d170cc16135783cULL, 0x3ff2c00000000000ULL, 0xbfc44d2b6ccb8000ULL,
    0x3d170cc16135783cULL, 0x3ff2a00000000000ULL, 0xbfc371fc201e9000ULL, 0x3cf178864d27543aULL,
    0x3ff2800000000000ULL, 0xbfc29552f81ff000ULL, 0xbd248d301771c408ULL, 0x3ff2600000000000ULL,
    0xbfc1b72ad52f6000ULL, 0xbd2e80a41811a396ULL, 0x3ff2600000000000ULL, 0xbfc1b72ad52f6000ULL,
    0xbd2e80a41811a396ULL, 0x3ff2400000000000ULL, 0xbfc0d77e7cd09000ULL, 0x3d0a699688e85bf4ULL,
    0x3ff2400000000000ULL, 0xbfc0d77e7cd09000ULL, 0x3d0a699688e85bf4ULL, 0x3ff2200000000000ULL,
    0xbfbfec9131dbe000ULL, 0xbd2575545ca333f2ULL, 0x3ff2000000000000ULL, 0xbfbe27076e2b0000ULL,
    0x3d2a342c2af0003cULL, 0x3ff2000000000000ULL, 0xbfbe27076e2b0000ULL, 0x3d2a342c2af0003cULL,
    0x3ff1e00000000000ULL, 0xbfbc5e548f5bc000ULL, 0xbd1d0c57585fbe06ULL, 0x3ff1c00000000000ULL,
    0xbfba926d3a4ae000ULL, 0x3d253935e85baac8ULL, 0x3ff1c00000000000ULL, 0xbfba926d3a4ae000ULL,
    0x3d253935e85baac8ULL, 0x3ff1a00000000000ULL, 0xbfb8c345d631a000ULL, 0x3d137c294d2f5668ULL,
    0x3ff1a00000000000ULL, 0xbfb8c345d631a000ULL, 0x3d137c294d2f5668ULL, 0x3ff1800000000000ULL,
    0xbfb6f0d28ae56000ULL, 0xbd269737c93373daULL, 0x3ff1600000000000ULL, 0xbfb51b073f062000ULL,
    0x3d1f025b61c65e57ULL, 0x3ff1600000000000ULL, 0xbfb51b073f062000ULL, 0x3d1f025b61c65e57ULL,
    0x3ff1400000000000ULL, 0xbfb341d7961be000ULL, 0x3d2c5edaccf913dfULL, 0x3ff1400000000000ULL,
    0xbfb341d7961be000ULL, 0x3d2c5edaccf913dfULL, 0x3ff1200000000000ULL, 0xbfb16536eea38000ULL,
    0x3d147c5e768fa309ULL, 0x3ff1000000000000ULL, 0xbfaf0a30c0118000ULL, 0x3d2d599e83368e91ULL,
    0x3ff1000000000000ULL, 0xbfaf0a30c0118000ULL, 0x3d2d599e83368e91ULL, 0x3ff0e00000000000ULL,
    0xbfab42dd71198000ULL, 0x3d1c827ae5d6704cULL, 0x3ff0e00000000000ULL, 0xbfab42dd71198000ULL,
    0x3d1c827ae5d6704cULL, 0x3ff0c00000000000ULL, 0xbfa77458f632c000ULL, 0xbd2cfc4634f2a1eeULL,
    0x3ff0c00000000000ULL, 0xbfa77458f632c000ULL, 0xbd2cfc4634f2a1eeULL, 0x3ff0a00000000000ULL,
    0xbfa39e87b9fec000ULL, 0x3cf502b7f526feaaULL, 0x3ff0a00000000000ULL, 0xbfa39e87b9fec000ULL,
    0x3cf502b7f526feaaULL, 0x3ff0800000000000ULL, 0xbf9f829b0e780000ULL, 0xbd2980267c7e09e4ULL,
    0x3ff0800000000000ULL, 0xbf9f829b0e780000ULL, 0xbd2980267c7e09e4ULL, 0x3ff0600000000000ULL,
    0xbf97b91b07d58000ULL, 0xbd288d5493faa639ULL, 0x3ff0400000000000ULL, 0xbf8fc0a8b0fc0000ULL,
    0xbcdf1e7cf6d3a69cULL, 0x3ff0400000000000ULL, 0xbf8fc0a8b0fc0000ULL, 0xbcdf1e7cf6d3a69cULL,
    0x3ff0200000000000ULL, 0xbf7fe02a6b100000ULL, 0xbd19e23f0dda40e4ULL, 0x3ff0200000000000ULL,
    0xbf7fe02a6b100000ULL, 0xbd19e23f0dda40e4ULL, 0x3ff0000000000000ULL, 0x0000000000000000ULL,
    0x0000000000000000ULL, 0x3ff0000000000000ULL, 0x0000000000000000ULL, 0x0000000000000000ULL,
    0x3fefc00000000000ULL, 0x3f80101575890000ULL, 0xbd10c76b999d2be8ULL, 0x3fef800000000000ULL,
    0x3f90205658938000ULL, 0xbd23dc5b06e2f7d2ULL, 0x3fef400000000000ULL, 0x3f98492528c90000ULL,
    0xbd2aa0ba325a0c34ULL, 0x3fef000000000000ULL, 0x3fa0415d89e74000ULL, 0x3d0111c05cf1d753ULL,
    0x3feec00000000000ULL, 0x3fa466aed42e0000ULL, 0xbd2c167375bdfd28ULL, 0x3fee800000000000ULL,
    0x3fa894aa149fc000ULL, 0xbd197995d05a267dULL, 0x3fee400000000000ULL, 0x3faccb73cdddc000ULL,
    0xbd1a68f247d82807ULL, 0x3fee200000000000ULL, 0x3faeea31c006c000ULL, 0xbd0e113e4fc93b7bULL,
    0x3fede00000000000ULL, 0x3fb1973bd1466000ULL, 0xbd25325d560d9e9bULL, 0x3feda00000000000ULL,
    0x3fb3bdf5a7d1e000ULL, 0x3d2cc85ea5db4ed7ULL, 0x3fed600000000000ULL, 0x3fb5e95a4d97a000ULL,
    0xbd2c69063c5d1d1eULL, 0x3fed400000000000ULL, 0x3fb700d30aeac000ULL, 0x3cec1e8da99ded32ULL,
    0x3fed000000000000ULL, 0x3fb9335e5d594000ULL, 0x3d23115c3abd47daULL, 0x3fecc00000000000ULL,
    0x3fbb6ac88dad6000ULL, 0xbd1390802bf768e5ULL, 0x3feca00000000000ULL, 0x3fbc885801bc4000ULL,
    0x3d2646d1c65aacd3ULL, 0x3fec600000000000ULL, 0x3fbec739830a2000ULL, 0xbd2dc068afe645e0ULL,
    0x3fec400000000000ULL, 0x3fbfe89139dbe000ULL, 0xbd2534d64fa10afdULL, 0x3fec000000000000ULL,
    0x3fc1178e8227e000ULL, 0x3d21ef78ce2d07f2ULL, 0x3febe00000000000ULL, 0x3fc1aa2b7e23f000ULL,
    0x3d2ca78e44389934ULL, 0x3feba00000000000ULL, 0x3fc2d1610c868000ULL, 0x3d039d6ccb81b4a1ULL,
    0x3feb800000000000ULL, 0x3fc365fcb0159000ULL, 0x3cc62fa8234b7289ULL, 0x3feb400000000000ULL,
    0x3fc4913d8333b000ULL, 0x3d25837954fdb678ULL, 0x3feb200000000000ULL, 0x3fc527e5e4a1b000ULL,
    0x3d2633e8e5697dc7ULL, 0x3feae00000000000ULL, 0x3fc6574ebe8c1000ULL, 0x3d19cf8b2c3c2e78ULL,
    0x3feac00000000000ULL, 0x3fc6f0128b757000ULL, 0xbd25118de59c21e1ULL, 0x3feaa00000000000ULL,
    0x3fc7898d85445000ULL, 0xbd1c661070914305ULL, 0x3fea600000000000ULL, 0x3fc8beafeb390000ULL,
    0xbd073d54aae92cd1ULL, 0x3fea400000000000ULL, 0x3fc95a5adcf70000ULL, 0x3d07f22858a0ff6fULL,
    0x3fea000000000000ULL, 0x3fca93ed3c8ae000ULL, 0xbd28724350562169ULL, 0x3fe9e00000000000ULL,
    0x3fcb31d8575bd000ULL, 0xbd0c358d4eace1aaULL, 0x3fe9c00000000000ULL, 0x3fcbd087383be000ULL,
    0xbd2d4bc4595412b6ULL, 0x3fe9a00000000000ULL, 0x3fcc6ffbc6f01000ULL, 0xbcf1ec72c5962bd2ULL,
    0x3fe9600000000000ULL, 0x3fcdb13db0d49000ULL, 0xbd2aff2af715b035ULL, 0x3fe9400000000000ULL,
    0x3fce530effe71000ULL, 0x3cc212276041f430ULL, 0x3fe9200000000000ULL, 0x3fcef5ade4dd0000ULL,
    0xbcca211565bb8e11ULL, 0x3fe9000000000000ULL, 0x3fcf991c6cb3b000ULL, 0x3d1bcbecca0cdf30ULL,
    0x3fe8c00000000000ULL, 0x3fd07138604d5800ULL, 0x3cf89cdb16ed4e91ULL, 0x3fe8a00000000000ULL,
    0x3fd0c42d67616000ULL, 0x3d27188b163ceae9ULL, 0x3fe8800000000000ULL, 0x3fd1178e8227e800ULL,
    0xbd2c210e63a5f01cULL, 0x3fe8600000000000ULL, 0x3fd16b5ccbacf800ULL, 0x3d2b9acdf7a51681ULL,
    0x3fe8400000000000ULL, 0x3fd1bf99635a6800ULL, 0x3d2ca6ed5147bdb7ULL, 0x3fe8200000000000ULL,
    0x3fd214456d0eb800ULL, 0x3d0a87deba46baeaULL, 0x3fe7e00000000000ULL, 0x3fd2bef07cdc9000ULL,
    0x3d2a9cfa4a5004f4ULL, 0x3fe7c00000000000ULL, 0x3fd314f1e1d36000ULL, 0xbd28e27ad3213cb8ULL,
    0x3fe7a00000000000ULL, 0x3fd36b6776be1000ULL, 0x3d116ecdb0f177c8ULL, 0x3fe7800000000000ULL,
    0x3fd3c25277333000ULL, 0x3d183b54b606bd5cULL, 0x3fe7600000000000ULL, 0x3fd419b423d5e800ULL,
    0x3d08e436ec90e09dULL, 0x3fe7400000000000ULL, 0x3fd4718dc271c800ULL, 0xbd2f27ce0967d675ULL,
    0x3fe7200000000000ULL, 0x3fd4c9e09e173000ULL, 0xbd2e20891b0ad8a4ULL, 0x3fe7000000000000ULL,
    0x3fd522ae0738a000ULL, 0x3d2ebe708164c759ULL, 0x3fe6e00000000000ULL, 0x3fd57bf753c8d000ULL,
    0x3d1fadedee5d40efULL, 0x3fe6c00000000000ULL, 0x3fd5d5bddf596000ULL, 0xbd0a0b2a08a465dcULL,
};

NGB_MATH_FN double ngb_pow(double x, double y)
{
    const double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    const double A0 = -0x1.0000000000000p-1, A1 = -0x1.5555555555560p-1, A2 = 0x1.0000000000006p-1, A3 = 0x1.999999959554ep-1,
                 A4 = -0x1.555555529a47ap-1, A5 = -0x1.2495b9b4845e9p+0, A6 = 0x1.0002b8b263fc3p+0;
    const double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8000000000000p+52;
    const double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    unsigned long long ix = ngb_d2bits(x);
    const unsigned long long iy = ngb_d2bits(y);
    unsigned topx = (unsigned)(ix >> 52);
    const unsigned topy = (unsigned)(iy >> 52);
    unsigned long long sign_bias = 0;
    if ((topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) return pow(x, y);
    if (topx >= 0x800u) {
        /* negative base (e_pow.c:305-318): an integer exponent continues on |x|, odd ones flip the sign */
        const int e = (int)(topy & 0x7ffu);
        int yint;
        if ((topx & 0x7ffu) - 0x001u >= 0x7ffu - 0x001u) return pow(x, y);
        if (e < 0x3ff) yint = 0;
        else if (e > 0x3ff + 52) yint = 2;
        else if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) yint = 0;
        else if (iy & (1ULL << (0x3ff + 52 - e))) yint = 1;
        else yint = 2;
        if (yint == 0) return pow(x, y);                     /* NaN */
        if (yint == 1) sign_bias = 0x800ULL << 7;
        ix &= 0x7fffffffffffffffULL;
        topx &= 0x7ffu;
    }
    if (topx - 0x001u >= 0x7ffu - 0x001u) return pow(x, y);
    /* log_inline */
    const unsigned long long tmp = ix - 0x3fe6955500000000ULL;
    const unsigned i = (unsigned)(tmp >> 45) & 127u;
    const int k = (int)((long long)tmp >> 52);
    const unsigned long long iz = ix - (tmp & 0xfff0000000000000ULL);
    const double z = ngb_bits2d(iz), kd = (double)k;
    const double invc = ngb_bits2d(NGB_TAB(ngb_pow_tab, 3 * i)), logc = ngb_bits2d(NGB_TAB(ngb_pow_tab, 3 * i + 1)),
                 logctail = ngb_bits2d(NGB_TAB(ngb_pow_tab, 3 * i + 2));
    const double r = fma(z, invc, -1.0);
    const double t1 = fma(kd, Ln2hi, logc);
    const double t2 = t1 + r;
    const double lo1 = fma(kd, Ln2lo, logctail);
    const double lo2 = (t1 - t2) + r;
    const double ar = A0 * r, ar2 = r * ar, ar3 = r * ar2;
    const double hi = t2 + ar2;
    const double lo3 = fma(ar, r, -ar2);
    const double lo4 = (t2 - hi) + ar2;
    const double pp = fma(ar2, fma(fma(r, A6, A5), ar2, fma(r, A4, A3)), fma(r, A2, A1));
    const double lo = fma(ar3, pp, ((lo1 + lo2) + lo3) + lo4);
    const double lhi = hi + lo;
    const double ltail = (hi - lhi) + lo;
    /* y * log(x) as ehi + elo */
    const double ehi = y * lhi;
    const double elo = fma(y, ltail, fma(lhi, y, -ehi));
    /* exp_inline(ehi, elo) */
    const unsigned abstop = (unsigned)(ngb_d2bits(ehi) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) return pow(x, y);
    double kq = fma(ehi, InvLn2N, Shift);
    const unsigned long long ki = ngb_d2bits(kq);
    kq -= Shift;
    double rr = fma(kq, NegLn2loN, fma(kq, NegLn2hiN, ehi));
    rr = elo + rr;
    const unsigned idx = 2u * (unsigned)(ki & 127u);
    const unsigned long long top = (ki + sign_bias) << 45;
    const double tail = ngb_bits2d(NGB_TAB(ngb_exp_tab, idx));
    const unsigned long long sbits = NGB_TAB(ngb_exp_tab, idx + 1) + top;
    const double r2 = rr * rr;
    const double p1 = fma(rr, C3, C2), p2 = fma(rr, C5, C4);
    const double tm = fma(r2 * r2, p2, fma(r2, p1, tail + rr));
    const double scale = ngb_bits2d(sbits);
    return fma(scale, tm, scale);
}

#endif
