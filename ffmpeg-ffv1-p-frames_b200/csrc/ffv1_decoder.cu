// ffv1_decoder.cu -- C-ABI decoder entry points (placeholder until the decode kernels land).
#include "../../include/ffv1_b200.h"
#include "ffv1_internal.h"
extern "C" {
int ffv1b200_dec_open(FFV1B200Decoder **dec, const FFV1B200DecParams *) { if (dec) *dec = nullptr; ffv1::set_last_error("decoder not built yet"); return FFV1B200_ERR_ENOSYS; }
void ffv1b200_dec_close(FFV1B200Decoder *) {}
int ffv1b200_dec_info(const FFV1B200Decoder *, FFV1B200DecInfo *) { return FFV1B200_ERR_ENOSYS; }
int ffv1b200_dec_decode_host(FFV1B200Decoder *, int, const uint8_t *const *, const int *, uint8_t *, size_t, int *, uint64_t *) { return FFV1B200_ERR_ENOSYS; }
int ffv1b200_dec_stats(const FFV1B200Decoder *, FFV1B200DecStats *) { return FFV1B200_ERR_ENOSYS; }
}
