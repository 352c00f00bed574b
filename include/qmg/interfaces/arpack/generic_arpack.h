// quantum-mg on B200 -- ARPACK is not part of the hot path: build with -DNO_ARPACK.
#ifndef QMG_B200_ARPACK
#define QMG_B200_ARPACK
#ifndef NO_ARPACK
#define NO_ARPACK
#endif
#endif
