// frontend_ref.cpp — adapter around the reference's UNCHANGED host text frontend.
//
// Compiled only inside/against the reference tree (needs its headers: processor/processor.h,
// cppjieba/Jieba.hpp, hanzi2phoneid.h, EnglishText2Id.h).  It walks the frontend tail of the `.bin`
// in the order the reference constructor does (src/models/SynthesizerTrn.cpp:169-298):
//   ENG : EnglishText2Id(blob + tail)                                                  (:169-176)
//   CHS : [tagger bytes, verbalizer bytes] FST pair -> wetext::Processor               (:180-207)
//         [dict, hmm, user, idf, stopword] -> cppjieba::Jieba                           (:209-271)
//         [polyphone words, polyphone pinyin] -> hanzi2phoneid                          (:273-298)
// Each section is preceded by its byte sizes stored as floats; after a section the byte cursor is
// re-aligned with the reference's own rule `off += off % 4` (:192-195) — kept as is.
#include <iostream>
#include <streambuf>

#include "EnglishText2Id.h"
#include "cppjieba/Jieba.hpp"
#include "frontend.hpp"
#include "hanzi2phoneid.h"
#include "processor/processor.h"

namespace {

struct MemBuf : std::streambuf {
    MemBuf(char* b, char* e) { setg(b, b, e); }
};

// Byte cursor over the float blob with the reference's alignment quirk.
struct Tail {
    float* base;
    int64_t bytes;
    int64_t off;       // float index of the next size header
    int64_t off_char;  // byte position after the last section
    bool more() const { return off_char + 1 < bytes; }
    int32_t size_hdr() { return (int32_t)base[off++]; }
    char* here() { return (char*)(base + off); }
    void skip(int64_t n) {
        off_char = off * 4 + n;
        if (off_char % 4 > 0) off_char += off_char % 4;
        off = off_char / 4;
    }
};

struct ChsFrontend : stts::Frontend {
    wetext::Processor* tn = nullptr;
    cppjieba::Jieba* jieba = nullptr;
    hanzi2phoneid* hz = nullptr;
    std::vector<std::string> words;
    ~ChsFrontend() { delete tn; delete jieba; delete hz; }
    bool text_to_ids(const std::string& line, std::vector<int32_t>& ids, float&) override {
        if (!jieba || !hz) return false;
        std::string s = line;                       // SynthesizerTrn.cpp:331-341
        if (tn) s = tn->verbalize(tn->tag(line));
        jieba->Cut(s, words, true);
        int32_t n = 0;
        int32_t* p = hz->convert(s, n, words);
        if (!p || n <= 0) return false;
        ids.assign(p, p + n);
        delete[] p;
        return true;
    }
};

struct EngFrontend : stts::Frontend {
    EnglishText2Id* g2p = nullptr;
    ~EngFrontend() { delete g2p; }
    bool text_to_ids(const std::string& line, std::vector<int32_t>& ids, float& ls) override {
        if (!g2p) return false;
        std::vector<int> v = g2p->getIPAId(line);   // SynthesizerTrn.cpp:343-355
        ids.assign(v.begin(), v.end());
        ls = ls * 0.83;
        return !ids.empty();
    }
};

}  // namespace

namespace stts {

Frontend* make_frontend(int32_t lang, float* data, int64_t bytes, int64_t tail_off) {
    if (lang == 1) {
        if (bytes <= (tail_off + 1) * 4) return nullptr;
        EngFrontend* f = new EngFrontend();
        int32_t cur = 0;
        f->g2p = new EnglishText2Id(data + tail_off, cur);
        return f;
    }
    Tail t{data, bytes, tail_off, tail_off * 4};
    ChsFrontend* f = new ChsFrontend();
    if (tail_off * 4 + 1 < bytes) {
        const int32_t a = t.size_hdr(), b = t.size_hdr();
        MemBuf tag(t.here(), t.here() + a), ver(t.here() + a, t.here() + a + b);
        std::istream it(&tag), iv(&ver);
        f->tn = new wetext::Processor(it, iv);
        t.skip((int64_t)a + b);
    }
    if (t.more()) {
        int32_t n[5];
        for (int i = 0; i < 5; ++i) n[i] = t.size_hdr();
        char* p = t.here();
        MemBuf b0(p, p + n[0]);               p += n[0];
        MemBuf b1(p, p + n[1]);               p += n[1];
        MemBuf b2(p, p + n[2]);               p += n[2];
        MemBuf b3(p, p + n[3]);               p += n[3];
        MemBuf b4(p, p + n[4]);
        std::istream i0(&b0), i1(&b1), i2(&b2), i3(&b3), i4(&b4);
        f->jieba = new cppjieba::Jieba(i0, i1, i2, i3, i4);
        t.skip((int64_t)n[0] + n[1] + n[2] + n[3] + n[4]);
    }
    if (t.more()) {
        const int32_t a = t.size_hdr(), b = t.size_hdr();
        MemBuf w(t.here(), t.here() + a), py(t.here() + a, t.here() + a + b);
        std::istream iw(&w), ip(&py);
        f->hz = new hanzi2phoneid(iw, ip);
        t.skip((int64_t)a + b);
    }
    return f;
}

}  // namespace stts
