// wav.hpp — the 16 kHz mono PCM16 RIFF/WAVE container the reference's demo writes
// (convertAudioToWavBuf, test/main.cpp:7-65): canonical 44-byte header, 'fmt ' chunk of 16 bytes, format 1,
// 1 channel, 16 000 Hz, byte rate 32 000, block align 2, 16 bits, then the 'data' chunk.  Same bytes for the same PCM.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace stts {

constexpr int kWavHeaderBytes = 44;
constexpr int kSampleRate = 16000;

inline void put_le32(uint8_t* p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = (v >> 24) & 0xff; }
inline void put_le16(uint8_t* p, uint16_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; }

// header for `data_bytes` bytes of PCM16 mono at `rate` Hz
inline void wav_header(uint8_t h[kWavHeaderBytes], uint32_t data_bytes, uint32_t rate = kSampleRate) {
    memcpy(h, "RIFF", 4);
    put_le32(h + 4, data_bytes + 36);          // RIFF chunk size = file size - 8
    memcpy(h + 8, "WAVEfmt ", 8);
    put_le32(h + 16, 16);                      // 'fmt ' chunk size
    put_le16(h + 20, 1);                       // PCM
    put_le16(h + 22, 1);                       // mono
    put_le32(h + 24, rate);
    put_le32(h + 28, rate * 2);                // byte rate
    put_le16(h + 32, 2);                       // block align
    put_le16(h + 34, 16);                      // bits per sample
    memcpy(h + 36, "data", 4);
    put_le32(h + 40, data_bytes);
}

// writes header + samples; returns false on I/O failure
inline bool write_wav(const std::string& path, const int16_t* pcm, size_t n_samples, uint32_t rate = kSampleRate) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    uint8_t h[kWavHeaderBytes];
    wav_header(h, (uint32_t)(n_samples * 2), rate);
    bool ok = fwrite(h, 1, kWavHeaderBytes, f) == (size_t)kWavHeaderBytes;
    if (ok && n_samples) ok = fwrite(pcm, 2, n_samples, f) == n_samples;   // little-endian host (x86-64 / aarch64)
    return fclose(f) == 0 && ok;
}

}  // namespace stts
