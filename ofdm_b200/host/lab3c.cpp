// lab3c.cpp -- the reference's lab3c example (examples/lab3c.rs:15-74) on the engine: --transmit writes an fc32 file,
// --receive reads one (optionally sliced with --start/--stop) and decodes it.
//   lab3c --transmit tx.dat --payload in.bin [--qpsk|--qam] [--guard]
//   lab3c --receive rx.dat --out out.bin [--start N] [--stop M] [--qpsk|--qam] [--guard] [--ecc]
// --ecc wraps the payload in the RS(255,223) outer code like examples/lab3c_image.rs:19-21,33-36.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>

#include "ofdm.hpp"

static std::vector<uint8_t> read_file(const std::string &p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", p.c_str()); std::exit(2); }
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}
static void write_file(const std::string &p, const std::vector<uint8_t> &b)
{
    std::ofstream f(p, std::ios::binary);
    f.write(reinterpret_cast<const char *>(b.data()), (std::streamsize)b.size());
}

int main(int argc, char **argv)
{
    std::string transmit, receive, payload, out;
    size_t start = 0, stop = (size_t)-1;
    bool guard = false, ecc = false;
    ofdm::ModulationScheme mod = ofdm::ModulationScheme::Bpsk;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return std::string(argv[++i]); };
        if (a == "--transmit") transmit = next();
        else if (a == "--receive") receive = next();
        else if (a == "--payload") payload = next();
        else if (a == "--out") out = next();
        else if (a == "--start") start = std::stoull(next());
        else if (a == "--stop") stop = std::stoull(next());
        else if (a == "--guard") guard = true;
        else if (a == "--ecc") ecc = true;
        else if (a == "--qpsk") mod = ofdm::ModulationScheme::Qpsk;
        else if (a == "--qam") mod = ofdm::ModulationScheme::Qam;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (transmit.empty() == receive.empty()) {
        std::fprintf(stderr, "Not a valid argument combination, specify transmit or receive, but not both\n");   // examples/lab3c.rs:82
        return 2;
    }
    try {
        ofdm::Modem modem(guard, mod);
        if (!transmit.empty()) {
            auto bytes = read_file(payload);
            if (ecc) bytes = ofdm::create_transmission_bytes(modem, bytes);
            auto samples = modem.encode(bytes);
            write_file(transmit, ofdm::sig_to_bytes(samples));
            std::printf("wrote %zu samples\n", samples.size());
        } else {
            auto samples = ofdm::bytes_to_sig(read_file(receive));
            if (stop > samples.size()) stop = samples.size();
            ofdm::SignalVec slice(samples.begin() + (std::ptrdiff_t)start, samples.begin() + (std::ptrdiff_t)stop);
            auto data = modem.decode(slice);
            if (ecc) {
                auto fixed = ofdm::decipher_transmission_bytes(modem, data);
                if (!fixed) throw std::runtime_error("Reed-Solomon: block beyond repair");
                data = *fixed;
            }
            write_file(out, data);
            std::printf("decoded %zu bytes\n", data.size());
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
