/* Hand-written config.h for compiling the reference ngspice sources in place
 * (test infrastructure only; autotools are not available in this image).
 * KLU on; XSPICE/CIDER/OSDI/PREDICTOR/NOBYPASS off; USE_OMP comes from -D on
 * the command line for the OpenMP flavour. */
#ifndef NGB200_REF_CONFIG_H
#define NGB200_REF_CONFIG_H
#define PACKAGE "ngspice"
#define PACKAGE_NAME "ngspice"
#define PACKAGE_TARNAME "ngspice"
#define PACKAGE_VERSION "45+"
#define PACKAGE_STRING "ngspice 45+"
#define PACKAGE_BUGREPORT "none"
#define PACKAGE_URL ""
#define VERSION "45+"
#define NGSPICEDATADIR "/nonexistent/share/ngspice"
#define NGSPICEBINDIR "/nonexistent/bin"
#define NGSPICEBUILDDATE "oracle"
#define KLU
#define SIMINFO
#define HAVE_STRING_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STDLIB_H 1
#define HAVE_UNISTD_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_TIME_H 1
#define HAVE_CTYPE_H 1
#define HAVE_FLOAT_H 1
#define HAVE_LIMITS_H 1
#define HAVE_MATH_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDBOOL_H 1
#define HAVE_DECL_ISNAN 1
#define HAVE_DECL_ISINF 1
#define HAVE_ISNAN 1
#define HAVE_ISINF 1
#define HAVE_FINITE 1
#define HAVE_LOGB 1
#define HAVE_SCALB 1
#define HAVE_SCALBN 1
#define HAVE_STRDUP 1
#define HAVE_STRNCASECMP 1
#define HAVE_ACCESS 1
#define HAVE_GETCWD 1
#define HAVE_GETPWUID 1
#define HAVE_PWD_H 1
#define HAVE_DIRENT_H 1
#define HAVE_GETRUSAGE 1
#define HAVE_SYS_RESOURCE_H 1
#define HAVE_TERMIOS_H 1
#define HAVE_ISATTY 1
#define HAVE_MEMMOVE 1
#define HAVE_MEMSET 1
#define HAVE_QSORT 1
#define HAVE_SNPRINTF 1
#define HAVE_VSNPRINTF 1
#define HAVE_POPEN 1
#define HAVE_DUP2 1
#define HAVE_FCNTL_H 1
#define STDC_HEADERS 1
#define RETSIGTYPE void
#define X_DISPLAY_MISSING 1
#endif
