// frontend.hpp — seam between the drop-in SynthesizerTrn shim and the HOST text frontend.
// The frontend itself is not part of this repository's scope (BASELINE.json: "the text frontend ...
// stays on the host CPU unchanged"); frontend_ref.cpp adapts the reference's own frontend objects.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace stts {

struct Frontend {
    virtual ~Frontend() {}
    // text -> phoneme ids; may scale length_scale (English: *0.83, SynthesizerTrn.cpp:354)
    virtual bool text_to_ids(const std::string& utf8, std::vector<int32_t>& ids, float& length_scale) = 0;
};

// Builds the frontend from the tail of the model blob (floats [tail_off, model_bytes/4)).
// Returns NULL when the tail is absent.  Implemented in frontend_ref.cpp (reference frontend).
Frontend* make_frontend(int32_t lang_type, float* model_data, int64_t model_bytes, int64_t tail_off);

}  // namespace stts
