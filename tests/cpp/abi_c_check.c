/* The public header is plain C: this file is compiled with `gcc -std=c99 -pedantic -c` by tests/test_abi.py. */
#include "../../include/jjschnorr_b200.h"

int jjs_abi_c_check(jjs_ctx* ctx, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status) {
    if (jjs_device_count(ctx) < 1) return JJS_ERR_ARGUMENT;
    return jjs_verify_single(ctx, pk, sig, msg, n, status, (uint8_t*)0) == JJS_SUCCESS ? JJS_OK : JJS_ERR_CUDA;
}
