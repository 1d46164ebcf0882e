// SynthesizerTrn_b200.cpp — drop-in replacement of the reference's src/models/SynthesizerTrn.cpp.
//
// Same class, same header (include/SynthesizerTrn.h:9-19 of the reference, compiled against
// unchanged), same ownership and error behaviour:
//   * ctor(float* modelData, int32_t modelSize /*bytes*/): the NN section is handed to
//     stts_create() (sm_100a kernels behind the C ABI, include/stts_b200.h); the frontend tail of the
//     blob (starting at stts_nn_end_offset) is handed to the HOST text frontend, which stays the
//     reference's own code (src/tn, src/hz2py, src/engipa, cppjieba — see frontend_ref.cpp).
//   * infer(line, sid, lengthScale, dataLen): text -> ids on the host (SynthesizerTrn.cpp:329-355,
//     incl. the English lengthScale*0.83 at :354), then ids -> PCM on the GPU (stts_infer_ids
//     replaces :357-396).  Returns a malloc'd int16 buffer the caller frees with tts_free_data.
//   * getSpeakerNum(): 0 -> 1 (:79-89).
// No exceptions escape; failures are logged through tts_log and leave the object inert
// (infer returns NULL, dataLen = 0) — the reference leaves priv_ dangling in that case (:94-98).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "SynthesizerTrn.h"   // the reference's public header, unchanged
#include "frontend.hpp"
#include "stts_b200.h"
#include "tts_logger.h"       // reference platform shim: tts_log(cat, str) == printf

namespace {
struct Priv {
    stts_engine* eng = nullptr;
    stts::Frontend* fe = nullptr;
    int32_t langType = 0;
    int32_t spkNum = 0;
};
void log_err(const char* what) {
    std::string s = std::string("SynthesizerTrn(b200): ") + what + ": " + stts_last_error() + "\n";
    tts_log(TTS_LOG_ERROR, s.c_str());
}
}  // namespace

SynthesizerTrn::SynthesizerTrn(float* modelData, int32_t modelSize) {
    Priv* P = new Priv();
    priv_ = P;
    int device = 0;
    if (const char* d = getenv("STTS_DEVICE")) device = atoi(d);
    if (stts_create(modelData, (int64_t)modelSize, device, &P->eng) != STTS_OK) {
        log_err("stts_create failed");
        P->eng = nullptr;
        return;
    }
    P->langType = stts_header_field(P->eng, 1);
    P->spkNum = stts_speaker_num(P->eng);
    const int64_t tail = stts_nn_end_offset(P->eng);  // floats
    P->fe = stts::make_frontend(P->langType, modelData, (int64_t)modelSize, tail);
}

int32_t SynthesizerTrn::getSpeakerNum() {
    Priv* P = (Priv*)priv_;
    return P->spkNum == 0 ? 1 : P->spkNum;
}

int16_t* SynthesizerTrn::infer(const string& line, int32_t sid, float lengthScale, int32_t& dataLen) {
    Priv* P = (Priv*)priv_;
    dataLen = 0;
    if (!P || !P->eng || !P->fe) return NULL;
    std::vector<int32_t> ids;
    float ls = lengthScale;
    if (!P->fe->text_to_ids(line, ids, ls)) {
        tts_log(TTS_LOG_ERROR, "SynthesizerTrn(b200): text frontend produced no ids\n");
        return NULL;
    }
    int16_t* pcm = NULL;
    int32_t n = 0;
    if (stts_infer_ids(P->eng, ids.data(), (int32_t)ids.size(), sid, ls, &pcm, &n) != STTS_OK) {
        log_err("stts_infer_ids failed");
        return NULL;
    }
    dataLen = n;
    return pcm;  // malloc'd: tts_free_data == free (src/utils/utils.cpp:34-37)
}

SynthesizerTrn::~SynthesizerTrn() {
    Priv* P = (Priv*)priv_;
    if (!P) return;
    delete P->fe;
    if (P->eng) stts_destroy(P->eng);
    delete P;
}
