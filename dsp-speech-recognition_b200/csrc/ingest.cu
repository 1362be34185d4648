// WAV ingest: RIFF/WAVE int16 files -> channel 0 -> the packed ragged PCM batch on the device.
// Replaces the per-file Python loop of reference reader.py:67-85 (scipy.io.wavfile.read + sig[:,0]).  The host only
// parses headers; sample bytes go to the device as they are, in slabs through two pinned staging buffers, and a
// kernel picks channel 0 out of the interleaved frames straight into its place in the packed buffer.
#include <cstdio>
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

// channel 0 of interleaved int16 frames -> packed destination, a whole slab of files per launch
struct FileDesc { int64_t src_off; int64_t dst_off; int64_t n_frames; int32_t channels; int32_t pad; };   // offsets in samples
__global__ void channel0_kernel(const int16_t* slab, const FileDesc* desc, int16_t* dst) {
    const FileDesc d = desc[blockIdx.y];
    const int16_t* src = slab + d.src_off;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n_frames; i += (int64_t)gridDim.x * blockDim.x)
        dst[d.dst_off + i] = src[i * d.channels];
}

// Staging buffers of the ingest path, allocated on first use and kept (pinned allocations cost tens of milliseconds).
struct IngestCtx {
    void* h_stage[2] = {nullptr, nullptr}; int16_t* d_stage[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
    FileDesc* h_desc[2] = {nullptr, nullptr}; FileDesc* d_desc[2] = {nullptr, nullptr};
    int64_t cap = 0; int device = -1;
    void release() {
        for (int i = 0; i < 2; ++i) {
            if (h_stage[i]) cudaFreeHost(h_stage[i]); if (d_stage[i]) cudaFree(d_stage[i]); if (ev[i]) cudaEventDestroy(ev[i]);
            if (h_desc[i]) cudaFreeHost(h_desc[i]); if (d_desc[i]) cudaFree(d_desc[i]);
            h_stage[i] = nullptr; d_stage[i] = nullptr; ev[i] = nullptr; h_desc[i] = nullptr; d_desc[i] = nullptr;
        }
        cap = 0; device = -1;
    }
};
IngestCtx g_ingest;
constexpr int kMaxFilesPerSlab = 4096;

int ingest_ensure(int64_t slab_cap) {
    int dev = 0; cudaGetDevice(&dev);
    if (g_ingest.cap >= slab_cap && g_ingest.device == dev) return DSPFE_OK;
    g_ingest.release();
    for (int i = 0; i < 2; ++i) {
        if (cudaHostAlloc(&g_ingest.h_stage[i], (size_t)slab_cap + 16, cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc(&g_ingest.d_stage[i], (size_t)slab_cap + 16) != cudaSuccess ||
            cudaEventCreateWithFlags(&g_ingest.ev[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaHostAlloc((void**)&g_ingest.h_desc[i], kMaxFilesPerSlab * sizeof(FileDesc), cudaHostAllocDefault) != cudaSuccess ||
            cudaMalloc(&g_ingest.d_desc[i], kMaxFilesPerSlab * sizeof(FileDesc)) != cudaSuccess) {
            g_ingest.release(); return fail(DSPFE_ERR_NOMEM, "staging allocation failed");
        }
    }
    g_ingest.cap = slab_cap; g_ingest.device = dev;
    return DSPFE_OK;
}

}  // namespace

extern "C" {

void dspfe_ingest_release(void) { g_ingest.release(); }


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
    // Files are packed into slabs of up to kSlabBytes of raw sample bytes; two pinned staging buffers and two device
    // buffers are ping-ponged, so the host fills slab s+1 while slab s crosses PCIe; one kernel launch per slab.
    const int64_t kSlabBytes = 32ll << 20;
    const int64_t slab_cap = max_bytes > kSlabBytes ? max_bytes : kSlabBytes;
    { const int rc = ingest_ensure(slab_cap); if (rc) return rc; }
    void** h_stage = g_ingest.h_stage; int16_t** d_stage = g_ingest.d_stage; cudaEvent_t* ev = g_ingest.ev;
    FileDesc** h_desc = g_ingest.h_desc; FileDesc** d_desc = g_ingest.d_desc;
    auto cleanup = [&]() {};
    int f = 0, slab = 0;
    while (f < n_files) {
        const int s = slab & 1;
        if (slab >= 2) cudaEventSynchronize(ev[s]);                      // the slot's previous slab has left the pinned buffers
        int64_t used = 0, longest = 0; int nd = 0;
        while (f < n_files && nd < kMaxFilesPerSlab) {
            const WavInfo& w = info[f];
            const int64_t nb = w.n_frames * 2 * w.channels;
            if (nd > 0 && used + nb > slab_cap) break;
            if (nb > 0) {
                std::memcpy((unsigned char*)h_stage[s] + used, (const unsigned char*)file_bytes[f] + w.data_offset, (size_t)nb);
                h_desc[s][nd++] = FileDesc{used / 2, h_offsets[f], w.n_frames, w.channels, 0};
                longest = w.n_frames > longest ? w.n_frames : longest;
                used += nb;
            }
            ++f;
        }
        if (nd == 0) { ++slab; continue; }
        cudaError_t e = cudaMemcpyAsync(d_stage[s], h_stage[s], (size_t)used, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc[s], h_desc[s], nd * sizeof(FileDesc), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            unsigned gx = (unsigned)((longest + 255) / 256); if (gx > 1024) gx = 1024; if (gx < 1) gx = 1;
            channel0_kernel<<<dim3(gx, (unsigned)nd), 256, 0, st>>>(d_stage[s], d_desc[s], d_pcm);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaEventRecord(ev[s], st);
        if (e != cudaSuccess) { cudaStreamSynchronize(st); cleanup(); return fail(DSPFE_ERR_CUDA, cudaGetErrorString(e)); }
        ++slab;
    }
    cudaStreamSynchronize(st);     // the batch is resident when the call returns; the staging buffers stay for the next call
    return DSPFE_OK;
}

/* ---- path-based ingest: the library reads the files itself, sample bytes go straight into the pinned slabs ---- */
// Walks the RIFF chunks of the file itself (8-byte chunk headers, seeks over the payloads): a `data` chunk behind a large
// LIST / bext / id3 chunk is found wherever it sits, as scipy.io.wavfile.read does (reader.py:76).
static int scan_one(const char* path, WavInfo& w, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return DSPFE_ERR_INVALID_ARG; }
    struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
    if (std::fseek(f, 0, SEEK_END) != 0) { err = "cannot seek"; return DSPFE_ERR_INVALID_ARG; }
    const long long fsize = std::ftell(f);
    if (fsize < 0 || std::fseek(f, 0, SEEK_SET) != 0) { err = "cannot seek"; return DSPFE_ERR_INVALID_ARG; }
    unsigned char h[48];
    if (std::fread(h, 1, 12, f) != 12 || std::memcmp(h, "RIFF", 4) != 0 || std::memcmp(h + 8, "WAVE", 4) != 0) { err = "not a RIFF/WAVE file"; return DSPFE_ERR_INVALID_ARG; }
    long long pos = 12;
    bool have_fmt = false;
    while (pos + 8 <= fsize) {
        if (std::fseek(f, (long)pos, SEEK_SET) != 0 || std::fread(h, 1, 8, f) != 8) break;
        const long long len = rd32(h + 4);
        if (std::memcmp(h, "fmt ", 4) == 0) {
            const size_t want = (size_t)(len < 40 ? len : 40);
            if (len < 16 || std::fread(h + 8, 1, want, f) != want) { err = "truncated fmt chunk"; return DSPFE_ERR_INVALID_ARG; }
            uint16_t tag = rd16(h + 8);
            w.channels = rd16(h + 10); w.rate = (int32_t)rd32(h + 12); w.bits = rd16(h + 22);
            if (tag == 0xFFFE && len >= 40) tag = rd16(h + 8 + 24);   // sub-format GUID starts with the tag
            if (tag != 1) { err = "only integer PCM WAV files are built"; return DSPFE_ERR_UNSUPPORTED; }
            if (w.bits != 16) { err = "only 16-bit WAV files are built (what the reference's data set holds)"; return DSPFE_ERR_UNSUPPORTED; }
            if (w.channels < 1) { err = "WAV file without channels"; return DSPFE_ERR_INVALID_ARG; }
            have_fmt = true;
        } else if (std::memcmp(h, "data", 4) == 0) {
            if (!have_fmt) { err = "data chunk before fmt chunk"; return DSPFE_ERR_INVALID_ARG; }
            long long n = len;
            if (pos + 8 + n > fsize) n = fsize - pos - 8;      // scipy also reads what is there
            w.data_offset = pos + 8;
            w.n_frames = n / (2 * w.channels);
            return DSPFE_OK;
        }
        pos += 8 + len + (len & 1);
    }
    err = "no data chunk";
    return DSPFE_ERR_INVALID_ARG;
}

int dspfe_wav_scan_paths(const char* const* paths, int32_t n_files, int64_t* h_offsets, int32_t* h_rates) {
    if (!paths || !h_offsets || n_files < 0) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    int64_t total = 0;
    for (int f = 0; f < n_files; ++f) {
        WavInfo w{}; std::string err;
        const int rc = scan_one(paths[f], w, err);
        if (rc) return fail(rc, "file " + std::to_string(f) + ": " + err);
        h_offsets[f] = total; total += w.n_frames;
        if (h_rates) h_rates[f] = w.rate;
    }
    h_offsets[n_files] = total;
    return DSPFE_OK;
}

int dspfe_ingest_wav_paths(const char* const* paths, int32_t n_files, int16_t* d_pcm, int64_t capacity, int64_t* h_offsets,
                           int32_t* h_rates, void* stream) {
    if (!paths || !h_offsets || n_files < 0 || (n_files > 0 && !d_pcm)) return fail(DSPFE_ERR_INVALID_ARG, "bad argument");
    std::vector<WavInfo> info(n_files);
    int64_t total = 0, max_bytes = 0;
    for (int f = 0; f < n_files; ++f) {
        std::string err;
        const int rc = scan_one(paths[f], info[f], err);
        if (rc) return fail(rc, "file " + std::to_string(f) + ": " + err);
        h_offsets[f] = total; total += info[f].n_frames;
        if (h_rates) h_rates[f] = info[f].rate;
        const int64_t nb = info[f].n_frames * 2 * info[f].channels;
        max_bytes = nb > max_bytes ? nb : max_bytes;
    }
    h_offsets[n_files] = total;
    if (total > capacity) return fail(DSPFE_ERR_INVALID_ARG, "d_pcm capacity is below the total sample count");
    if (n_files == 0 || total == 0) return DSPFE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t kSlabBytes = 32ll << 20;
    const int64_t slab_cap = max_bytes > kSlabBytes ? max_bytes : kSlabBytes;
    { const int rc = ingest_ensure(slab_cap); if (rc) return rc; }
    int f = 0, slab = 0;
    while (f < n_files) {
        const int s = slab & 1;
        if (slab >= 2) cudaEventSynchronize(g_ingest.ev[s]);
        int64_t used = 0, longest = 0; int nd = 0;
        while (f < n_files && nd < kMaxFilesPerSlab) {
            const WavInfo& w = info[f];
            const int64_t nb = w.n_frames * 2 * w.channels;
            if (nd > 0 && used + nb > slab_cap) break;
            if (nb > 0) {
                FILE* fp = std::fopen(paths[f], "rb");
                bool ok = fp != nullptr;
                if (ok) { std::fseek(fp, (long)w.data_offset, SEEK_SET); ok = std::fread((unsigned char*)g_ingest.h_stage[s] + used, 1, (size_t)nb, fp) == (size_t)nb; std::fclose(fp); }
                if (!ok) { cudaStreamSynchronize(st); return fail(DSPFE_ERR_INVALID_ARG, std::string("short read: ") + paths[f]); }
                g_ingest.h_desc[s][nd++] = FileDesc{used / 2, h_offsets[f], w.n_frames, w.channels, 0};
                longest = w.n_frames > longest ? w.n_frames : longest;
                used += nb;
            }
            ++f;
        }
        if (nd == 0) { ++slab; continue; }
        cudaError_t e = cudaMemcpyAsync(g_ingest.d_stage[s], g_ingest.h_stage[s], (size_t)used, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(g_ingest.d_desc[s], g_ingest.h_desc[s], nd * sizeof(FileDesc), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            unsigned gx = (unsigned)((longest + 255) / 256); if (gx > 1024) gx = 1024; if (gx < 1) gx = 1;
            channel0_kernel<<<dim3(gx, (unsigned)nd), 256, 0, st>>>(g_ingest.d_stage[s], g_ingest.d_desc[s], d_pcm);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaEventRecord(g_ingest.ev[s], st);
        if (e != cudaSuccess) { cudaStreamSynchronize(st); return fail(DSPFE_ERR_CUDA, cudaGetErrorString(e)); }
        ++slab;
    }
    cudaStreamSynchronize(st);
    return DSPFE_OK;
}

}  // extern "C"
