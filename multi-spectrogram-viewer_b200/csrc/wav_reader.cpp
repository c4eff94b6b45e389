// wav_reader.cpp -- host-side WAV decode standing in for audio::open_audio_file's hound branch
// (audio.rs:9-21).  Integer samples are scaled by 1/2^(bits-1) (audio.rs:16-19); 16-bit files are
// kept as int16 so the (exact) scaling can run on the GPU while loading.
#include <cerrno>
#include <cstdio>
#include <cstring>

#include <algorithm>

#include "engine.h"

namespace sgx {

namespace {
struct File {
    FILE *f = nullptr;
    explicit File(const std::string &p) { f = std::fopen(p.c_str(), "rb"); }
    ~File() { if (f) std::fclose(f); }
};
uint32_t rd32(const unsigned char *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }
uint16_t rd16(const unsigned char *b) { return (uint16_t)(b[0] | (b[1] << 8)); }
} // namespace

WavData read_wav(const std::string &path)
{
    File fh(path);
    if (!fh.f) throw Error(SGX_ERR_IO, path + ": " + std::strerror(errno));
    // chunk sizes come from the file: never trust them beyond what the file can hold (a truncated or streamed WAV
    // announces 0xFFFFFFFF, which would otherwise turn into a 4 GiB allocation)
    long file_size = -1;
    if (std::fseek(fh.f, 0, SEEK_END) == 0) file_size = std::ftell(fh.f);
    if (file_size < 0 || std::fseek(fh.f, 0, SEEK_SET) != 0) throw Error(SGX_ERR_IO, path + ": cannot determine the file size");
    auto remaining = [&]() -> size_t { const long at = std::ftell(fh.f); return at < 0 || at > file_size ? 0 : (size_t)(file_size - at); };
    unsigned char hdr[12];
    if (std::fread(hdr, 1, 12, fh.f) != 12 || std::memcmp(hdr, "RIFF", 4) != 0 || std::memcmp(hdr + 8, "WAVE", 4) != 0)
        throw Error(SGX_ERR_IO, path + ": not a RIFF/WAVE file (only the WAV branch of audio.rs:9-21 is supported)");
    uint16_t fmt_tag = 0, channels = 0, bits = 0, block_align = 0;
    uint32_t sr = 0;
    bool have_fmt = false;
    WavData out;
    for (;;) {
        unsigned char ch[8];
        if (std::fread(ch, 1, 8, fh.f) != 8) break;
        const uint32_t size = rd32(ch + 4);
        if (std::memcmp(ch, "fmt ", 4) == 0) {
            if (size < 16 || size > remaining() || size > 4096) throw Error(SGX_ERR_IO, path + ": truncated or oversized fmt chunk");
            std::vector<unsigned char> b(size);
            if (std::fread(b.data(), 1, size, fh.f) != size) throw Error(SGX_ERR_IO, path + ": truncated fmt chunk");
            fmt_tag = rd16(&b[0]); channels = rd16(&b[2]); sr = rd32(&b[4]); block_align = rd16(&b[12]); bits = rd16(&b[14]);
            if (fmt_tag == 0xFFFE && size >= 26) fmt_tag = rd16(&b[24]); // WAVE_FORMAT_EXTENSIBLE sub-format
            have_fmt = true;
            if ((size & 1) && std::fseek(fh.f, 1, SEEK_CUR) != 0) break;
        } else if (std::memcmp(ch, "data", 4) == 0) {
            if (!have_fmt) throw Error(SGX_ERR_IO, path + ": data chunk before fmt chunk");
            if (channels == 0 || sr == 0) throw Error(SGX_ERR_IO, path + ": invalid fmt chunk");
            const size_t bytes_per = bits / 8;
            if (bytes_per == 0 || block_align != bytes_per * channels) throw Error(SGX_ERR_IO, path + ": unsupported sample packing");
            const size_t take = std::min<size_t>(size, remaining()); // what the file really holds
            std::vector<unsigned char> raw(take);
            const size_t got = take ? std::fread(raw.data(), 1, take, fh.f) : 0;
            const size_t n = got / block_align;
            out.sr = sr; out.ch = channels; out.n = n;
            const size_t total = n * channels;
            if (fmt_tag == 1 && bits == 16) {
                out.is_i16 = true;
                out.i16.resize(total);
                for (size_t i = 0; i < total; ++i) out.i16[i] = (int16_t)rd16(&raw[2 * i]);
            } else if (fmt_tag == 1 && (bits == 8 || bits == 24 || bits == 32)) {
                out.f32.resize(total);
                const float scale = (float)(1u << (bits - 1)); // audio.rs:18  2^(bits-1)
                for (size_t i = 0; i < total; ++i) {
                    int32_t v;
                    if (bits == 8) v = (int32_t)raw[i] - 128; // hound: unsigned 8-bit -> signed
                    else if (bits == 24) { v = raw[3 * i] | (raw[3 * i + 1] << 8) | (raw[3 * i + 2] << 16); if (v & 0x800000) v |= ~0xFFFFFF; }
                    else v = (int32_t)rd32(&raw[4 * i]);
                    out.f32[i] = (float)v / scale;
                }
            } else if (fmt_tag == 3 && bits == 32) {
                out.f32.resize(total);
                std::memcpy(out.f32.data(), raw.data(), total * 4); // audio.rs:14 float passthrough
            } else {
                throw Error(SGX_ERR_IO, path + ": unsupported WAV encoding (tag " + std::to_string(fmt_tag) + ", " + std::to_string(bits) + " bits)");
            }
            return out;
        } else {
            const size_t skip = (size_t)size + (size & 1);
            if (skip > remaining() || std::fseek(fh.f, (long)skip, SEEK_CUR) != 0) break; // chunk runs past the end of the file
        }
    }
    throw Error(SGX_ERR_IO, path + ": no data chunk");
}

} // namespace sgx
