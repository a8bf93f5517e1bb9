// ofdm.hpp -- C++ host surface above the C ABI (include/ofdm_engine.h), mirroring the reference crate's public API
// for the modem path. The reference is compiled Rust code; without a Rust toolchain in the build image the same
// surface is provided in C++ (rust/ofdm-sys holds the equivalent Rust shim).
//
//   ofdm::encode(data, guard_bands, modulation)  -> std::vector<std::complex<double>>   src/transmitter.rs:10-58
//   ofdm::decode(samples, guard_bands, modulation) -> std::vector<uint8_t> (throws)     src/receiver.rs:8-96
//   ofdm::ModulationScheme {Bpsk, Qpsk, Qam}                                            src/transmitter.rs:98-104
//   ofdm::Header                                                                         src/packets/mod.rs:20-32
//   ofdm::Analysis                                                                       src/utils.rs:38-69
//   ofdm::sig_to_bytes / bytes_to_sig                                                    src/utils.rs:228-254
//   ofdm::create_transmission_bytes / decipher_transmission_bytes (RS(255,223))          src/utils.rs:97-137,152-180
//   ofdm::Modem::encode_batch / decode_batch : the batched forms a GPU needs
#pragma once

#include <complex>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ofdm_engine.h"

namespace ofdm {

enum class ModulationScheme : uint32_t { Bpsk = OFDM_MOD_BPSK, Qpsk = OFDM_MOD_QPSK, Qam = OFDM_MOD_QAM64 };

using Complex64 = std::complex<double>;
using SignalVec = std::vector<Complex64>;

// anyhow::Error of `decode` (src/receiver.rs:27-29) and its panics (:25, :87-89), one status per stream
struct DecodeError : std::runtime_error {
    int32_t status;
    explicit DecodeError(int32_t st)
        : std::runtime_error(st == OFDM_TOO_SHORT ? "Input not long enough, bailing early" : ofdm_status_name(st)), status(st) {}
};

struct Header {                       // bincode fixint: u128 little endian
    uint64_t packet_length = 0;
    std::vector<uint8_t> serialize() const
    {
        std::vector<uint8_t> b(16, 0);
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(packet_length >> (8 * i));
        return b;
    }
    static Header deserialize(const uint8_t *b)
    {
        Header h;
        for (int i = 0; i < 8; i++) h.packet_length |= (uint64_t)b[i] << (8 * i);
        return h;
    }
};

// fc32 wire format (src/utils.rs:228-254)
inline std::vector<uint8_t> sig_to_bytes(const SignalVec &sig)
{
    std::vector<uint8_t> out(sig.size() * 8);
    for (size_t i = 0; i < sig.size(); i++) {
        float re = (float)sig[i].real(), im = (float)sig[i].imag();
        std::memcpy(&out[8 * i], &re, 4);
        std::memcpy(&out[8 * i + 4], &im, 4);
    }
    return out;
}
inline SignalVec bytes_to_sig(const std::vector<uint8_t> &in)
{
    SignalVec out(in.size() / 8);
    for (size_t i = 0; i < out.size(); i++) {
        float re, im;
        std::memcpy(&re, &in[8 * i], 4);
        std::memcpy(&im, &in[8 * i + 4], 4);
        out[i] = Complex64(re, im);
    }
    return out;
}

// One engine handle (one GPU). Not thread-safe, like the C handle.
class Modem {
public:
    explicit Modem(std::optional<bool> guard_bands = std::nullopt, std::optional<ModulationScheme> modulation = std::nullopt,
                   bool fec = false, int device = 0, uint32_t sync_mode = OFDM_SYNC_REFERENCE, uint32_t cfo_mode = OFDM_CFO_REFERENCE,
                   uint32_t phase_mode = OFDM_PHASE_REFERENCE, uint32_t sync_window = 0)
    {
        ofdm_cfg_default(&cfg_);
        cfg_.guard_bands = guard_bands.value_or(false) ? 1 : 0;                                   // src/transmitter.rs:16
        cfg_.modulation = (uint32_t)modulation.value_or(ModulationScheme::Bpsk);                   // src/transmitter.rs:17
        cfg_.fec = fec ? 1 : 0;
        cfg_.sync_mode = sync_mode; cfg_.cfo_mode = cfo_mode; cfg_.phase_mode = phase_mode; cfg_.sync_window = sync_window;
        if (ofdm_engine_create(&cfg_, device, &h_) != 0) throw std::runtime_error(std::string("ofdm_engine_create: ") + ofdm_last_error(nullptr));
    }
    ~Modem() { ofdm_engine_destroy(h_); }
    Modem(const Modem &) = delete;
    Modem &operator=(const Modem &) = delete;

    std::vector<SignalVec> encode_batch(const std::vector<std::vector<uint8_t>> &payloads)
    {
        const uint32_t n = (uint32_t)payloads.size();
        uint32_t pstride = 1, istride = 880;
        std::vector<uint32_t> len(n), flen(n);
        for (uint32_t i = 0; i < n; i++) {
            len[i] = (uint32_t)payloads[i].size();
            if (len[i] > pstride) pstride = len[i];
            uint32_t f = ofdm_frame_len(&cfg_, len[i]);
            if (f > istride) istride = f;
        }
        std::vector<uint8_t> pay((size_t)n * pstride, 0);
        for (uint32_t i = 0; i < n; i++) std::memcpy(&pay[(size_t)i * pstride], payloads[i].data(), len[i]);
        std::vector<ofdm_fc32> iq((size_t)n * istride);
        check(ofdm_tx_encode_batch(h_, pay.data(), len.data(), pstride, n, iq.data(), istride, flen.data(), OFDM_MEM_HOST, nullptr));
        std::vector<SignalVec> out(n);
        for (uint32_t i = 0; i < n; i++) {
            out[i].resize(flen[i]);
            for (uint32_t k = 0; k < flen[i]; k++) out[i][k] = Complex64(iq[(size_t)i * istride + k].re, iq[(size_t)i * istride + k].im);
        }
        return out;
    }

    // returns the decoded payloads; status[i] != OFDM_OK marks a failed stream (its payload is empty)
    std::vector<std::vector<uint8_t>> decode_batch(const std::vector<SignalVec> &captures, std::vector<int32_t> &status)
    {
        const uint32_t n = (uint32_t)captures.size();
        uint32_t istride = 1;
        std::vector<uint32_t> ns(n), olen(n);
        for (uint32_t i = 0; i < n; i++) { ns[i] = (uint32_t)captures[i].size(); if (ns[i] > istride) istride = ns[i]; }
        std::vector<ofdm_fc32> iq((size_t)n * istride, ofdm_fc32{0, 0});
        for (uint32_t i = 0; i < n; i++)
            for (uint32_t k = 0; k < ns[i]; k++) iq[(size_t)i * istride + k] = ofdm_fc32{(float)captures[i][k].real(), (float)captures[i][k].imag()};
        const uint32_t ostride = (istride / 80 + 1) * 48 + 16;          // >= bytes of any frame that fits the capture
        std::vector<uint8_t> out((size_t)n * ostride);
        status.assign(n, 0);
        check(ofdm_rx_decode_batch(h_, iq.data(), ns.data(), n, istride, istride, out.data(), ostride, olen.data(), status.data(), nullptr,
                                   OFDM_MEM_HOST, nullptr));
        std::vector<std::vector<uint8_t>> res(n);
        for (uint32_t i = 0; i < n; i++)
            if (status[i] == OFDM_OK) res[i].assign(out.begin() + (size_t)i * ostride, out.begin() + (size_t)i * ostride + olen[i]);
        return res;
    }

    SignalVec encode(const std::vector<uint8_t> &data) { return encode_batch({data})[0]; }
    std::vector<uint8_t> decode(const SignalVec &samples)
    {
        std::vector<int32_t> st;
        auto r = decode_batch({samples}, st);
        if (st[0] != OFDM_OK) throw DecodeError(st[0]);
        return r[0];
    }
    ofdm_engine *handle() { return h_; }
    const ofdm_cfg &cfg() const { return cfg_; }

private:
    void check(int rc) { if (rc != 0) throw std::runtime_error(ofdm_last_error(h_)); }
    ofdm_cfg cfg_{};
    ofdm_engine *h_ = nullptr;
};

// the reference's free functions (one engine per call: fine for the lab examples, use Modem for throughput)
inline SignalVec encode(const std::vector<uint8_t> &data, std::optional<bool> guard_bands = std::nullopt,
                        std::optional<ModulationScheme> modulation = std::nullopt)
{
    return Modem(guard_bands, modulation).encode(data);
}
inline std::vector<uint8_t> decode(const SignalVec &samples, std::optional<bool> guard_bands = std::nullopt,
                                   std::optional<ModulationScheme> modulation = std::nullopt)
{
    return Modem(guard_bands, modulation).decode(samples);
}

// RS(255,223) outer code with the reference's block framing
inline std::vector<uint8_t> create_transmission_bytes(Modem &m, const std::vector<uint8_t> &data)      // src/utils.rs:97-137
{
    const uint32_t n = (uint32_t)data.size();
    std::vector<uint8_t> coded(ofdm_rs_encoded_len(n));
    uint32_t clen = 0;
    const uint8_t zero = 0;
    if (ofdm_rs_encode_batch(m.handle(), n ? data.data() : &zero, &n, 1, n ? n : 1, coded.data(), (uint32_t)coded.size(), &clen,
                             OFDM_MEM_HOST, nullptr) != 0)
        throw std::runtime_error(ofdm_last_error(m.handle()));
    coded.resize(clen);
    return coded;
}
inline std::optional<std::vector<uint8_t>> decipher_transmission_bytes(Modem &m, const std::vector<uint8_t> &coded)   // src/utils.rs:152-180
{
    const uint32_t n = (uint32_t)coded.size();
    std::vector<uint8_t> data(ofdm_rs_decoded_len(n));
    uint32_t dlen = 0, fixed = 0, failed = 0;
    const uint8_t zero = 0;
    if (ofdm_rs_decode_batch(m.handle(), n ? coded.data() : &zero, &n, 1, n ? n : 1, data.data(), (uint32_t)data.size(), &dlen, &fixed,
                             &failed, OFDM_MEM_HOST, nullptr) != 0)
        throw std::runtime_error(ofdm_last_error(m.handle()));
    if (failed) return std::nullopt;                                 // `.ok()?`, src/utils.rs:165,172
    data.resize(dlen);
    return data;
}

struct Analysis {                     // src/utils.rs:38-69
    uint32_t num_errs = 0, num_block_errs = 0;
    double err_rate = 0.0;
    static Analysis make(Modem &m, const std::vector<uint8_t> &left, const std::vector<uint8_t> &right)
    {
        if (left.size() != right.size()) throw std::invalid_argument("Analysis: length mismatch");     // assert_eq!, src/utils.rs:46
        uint32_t n = (uint32_t)left.size();
        int32_t st = 0;
        uint64_t c[4] = {0, 0, 0, 0};
        if (ofdm_ber_accumulate(m.handle(), left.data(), &n, n, right.data(), &n, n, &st, 1, c, OFDM_MEM_HOST, nullptr) != 0)
            throw std::runtime_error(ofdm_last_error(m.handle()));
        Analysis a;
        a.num_errs = (uint32_t)c[0]; a.num_block_errs = (uint32_t)c[1];
        a.err_rate = (double)c[0] / ((double)n * 8.0);
        return a;
    }
};

}  // namespace ofdm
