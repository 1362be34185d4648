// WAV ingest: RIFF/WAVE int16 files -> channel 0 -> the packed ragged PCM batch on the device.
// Replaces the per-file Python loop of reference reader.py:67-85 (scipy.io.wavfile.read + sig[:,0]).  The host only
// parses headers; sample bytes go to the device as they are, in slabs through two pinned staging buffers, and a
// kernel picks channel 0 out of the interleaved frames straight into its place in the packed buffer.
#include <cstring>
#include <vector>

#include "abi_common.h"

using namespace dspfe;

namespace {

struct WavInfo { int32_t rate, channels, bits; int64_t n_frames, data_offset; };

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// walks the RIFF chunks; accepts PCM (format 1) and WAVE_FORMAT_EXTENSIBLE with a PCM sub-format
int parse_wav(const unsigned char* b, int64_t size, WavInfo& w, std::string& err) {
    if (size < 12 || std::memcmp(b, "RIFF", 4) != 0 || std::memcmp(b + 8, "WAVE", 4) != 0) { err = "not a RIFF/WAVE file"; return DSPFE_ERR_INVALID_ARG; }
    int64_t pos = 12;
    bool have_fmt = false;
    while (pos + 8 <= size) {
        const unsigned char* c = b + pos;
        const int64_t len = rd32(c + 4);
        if (std::memcmp(c, "fmt ", 4) == 0) {
            if (len < 16 || pos + 8 + 16 > size) { err = "truncated fmt chunk"; return DSPFE_ERR_INVALID_ARG; }
            uint16_t tag = rd16(c + 8);
            w.channels = rd16(c + 10); w.rate = (int32_t)rd32(c + 12); w.bits = rd16(c + 22);
            if (tag == 0xFFFE && len >= 40 && pos + 8 + 26 <= size) tag = rd16(c + 8 + 24);   // sub-format GUID starts with the tag
            if (tag != 1) { err = "only integer PCM WAV files are built"; return DSPFE_ERR_UNSUPPORTED; }
            if (w.bits != 16) { err = "only 16-bit WAV files are built (what the reference's data set holds)"; return DSPFE_ERR_UNSUPPORTED; }
            if (w.channels < 1) { err = "WAV file without channels"; return DSPFE_ERR_INVALID_ARG; }
            have_fmt = true;
        } else if (std::memcmp(c, "data", 4) == 0) {
            if (!have_fmt) { err = "data chunk before fmt chunk"; return DSPFE_ERR_INVALID_ARG; }
            int64_t n = len;
            if (pos + 8 + n > size) n = size - pos - 8;      // scipy also reads what is there
            w.data_offset = pos + 8;
            w.n_frames = n / (2 * w.channels);
            return DSPFE_OK;
        }
        pos += 8 + len + (len & 1);
    }
    err = "no data chunk";
    return DSPFE_ERR_INVALID_ARG;
}

// channel 0 of interleaved int16 frames -> packed destination
__global__ void channel0_kernel(const int16_t* src, int64_t n_frames, int channels, int16_t* dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_frames) dst[i] = src[i * channels];
}

}  // namespace

extern "C" {

int dspfe_wav_info(const void* bytes, int64_t size, int32_t* rate, int32_t* channels, int32_t* bits, int64_t* n_frames, int64_t* data_offset) {
    if (!bytes || size < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    WavInfo w{}; std::string err;
    const int rc = parse_wav((const unsigned char*)bytes, size, w, err);
    if (rc) return fail(rc, err);
    if (rate) *rate = w.rate; if (channels) *channels = w.channels; if (bits) *bits = w.bits;
    if (n_frames) *n_frames = w.n_frames; if (data_offset) *data_offset = w.data_offset;
    return DSPFE_OK;
}

int dspfe_ingest_wavs(const void* const* file_bytes, const int64_t* sizes, int32_t n_files, int16_t* d_pcm, int64_t capacity,
                      int64_t* h_offsets, int32_t* h_rates, void* stream) {
    if (!file_bytes || !sizes || !h_offsets || n_files < 0 || (n_files > 0 && !d_pcm)) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    std::vector<WavInfo> info(n_files);
    int64_t total = 0, max_bytes = 0;
    for (int f = 0; f < n_files; ++f) {
        std::string err;
        const int rc = parse_wav((const unsigned char*)file_bytes[f], sizes[f], info[f], err);
        if (rc) return fail(rc, "file " + std::to_string(f) + ": " + err);
        h_offsets[f] = total;
        total += info[f].n_frames;
        if (h_rates) h_rates[f] = info[f].rate;
        const int64_t nb = info[f].n_frames * 2 * info[f].channels;
        max_bytes = nb > max_bytes ? nb : max_bytes;
    }
    h_offsets[n_files] = total;
    if (total > capacity) return fail(DSPFE_ERR_INVALID_ARG, "d_pcm capacity is below the total sample count");
    if (n_files == 0 || total == 0) return DSPFE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // two pinned staging buffers + two device buffers, ping-ponged: file f+1 is copied while file f's H2D is in flight
    void* h_stage[2] = {nullptr, nullptr}; int16_t* d_stage[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
    auto cleanup = [&]() { for (int i = 0; i < 2; ++i) { if (h_stage[i]) cudaFreeHost(h_stage[i]); if (d_stage[i]) cudaFree(d_stage[i]); if (ev[i]) cudaEventDestroy(ev[i]); } };
    for (int i = 0; i < 2; ++i) {
        if (cudaHostAlloc(&h_stage[i], (size_t)max_bytes + 16, cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc(&d_stage[i], (size_t)max_bytes + 16) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
            cleanup(); return fail(DSPFE_ERR_NOMEM, "staging allocation failed");
        }
    }
    for (int f = 0; f < n_files; ++f) {
        const int s = f & 1;
        const WavInfo& w = info[f];
        if (w.n_frames == 0) continue;
        const int64_t nb = w.n_frames * 2 * w.channels;
        if (f >= 2) cudaEventSynchronize(ev[s]);                         // the slot's previous file has left the pinned buffer
        std::memcpy(h_stage[s], (const unsigned char*)file_bytes[f] + w.data_offset, (size_t)nb);
        cudaError_t e = cudaMemcpyAsync(d_stage[s], h_stage[s], (size_t)nb, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            channel0_kernel<<<(unsigned)((w.n_frames + 255) / 256), 256, 0, st>>>(d_stage[s], w.n_frames, w.channels, d_pcm + h_offsets[f]);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaEventRecord(ev[s], st);
        if (e != cudaSuccess) { cudaStreamSynchronize(st); cleanup(); return fail(DSPFE_ERR_CUDA, cudaGetErrorString(e)); }
    }
    cudaStreamSynchronize(st);     // the staging buffers are released here; ingest is not on the kernels' hot path
    cleanup();
    return DSPFE_OK;
}

}  // extern "C"
