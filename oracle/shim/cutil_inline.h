/* oracle/shim/cutil_inline.h -- TEST INFRASTRUCTURE.  The reference includes <cutil_inline.h>
 * (photonMappingKernel.cu:5) but uses nothing from it; an empty stand-in is enough. */
