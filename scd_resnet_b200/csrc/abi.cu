// ABI bookkeeping for libscd_b200.so.
#include "common.cuh"

extern "C" int scd_abi_version(void) { return SCD_ABI_VERSION; }
extern "C" const char* scd_last_error(void) { return scd::err_buf(); }
