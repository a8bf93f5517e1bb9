"""C-ABI library loads and exports what include/ofdm_engine.h declares; host-side logic (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ofdm_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(engine_lib):
    from ofdm_b200 import engine
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(engine_lib, n), f"{n} declared in include/ofdm_engine.h but not exported"
    assert sorted(engine.EXPORTS) == names
    assert engine_lib.ofdm_abi_version() == 1


def test_no_torch_in_abi_signatures():
    text = open(os.path.join(ROOT, "include", "ofdm_engine.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "std::" not in text


def test_cfg_struct_layout(engine_lib):
    from ofdm_b200 import engine
    c = engine.CCfg()
    engine_lib.ofdm_cfg_default(C.byref(c))
    assert c.struct_size == C.sizeof(engine.CCfg) == 64
    assert (c.nfft, c.cp, c.modulation, c.guard_bands) == (64, 16, 0, 0)       # reference defaults
    assert engine_lib.ofdm_status_name(1) == b"TOO_SHORT"


def test_sizes_match_oracle(engine_lib, oo):
    import ofdm_b200 as ob
    for mod in (0, 1, 2):
        for guard in (False, True):
            for fec in (False, True):
                cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec)
                ocfg = oo.make_cfg(guard, mod, fec)
                for n in (0, 1, 8, 100, 576, 765, 41915):
                    assert cfg.frame_len(n) == oo.lib().oo_tx_len(n, C.byref(ocfg))
                    assert cfg.coded_len(n) == (oo.lib().oo_hamming74_encoded_len(n) if fec else n)
                for S in (1, 3, 17, 29, 131, 2038):
                    p = cfg.max_payload(S)
                    if p > 0:
                        assert cfg.frame_data_syms(p) <= S < cfg.frame_data_syms(p + 2) + 1
    assert ob.Config(modulation=2, guard_bands=True, fec=True).frame_len(576) == 3120


def test_engine_create_fails_loudly_without_gpu(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ofdm_b200 as ob
    with pytest.raises(ob.EngineError, match="no CPU fallback"):
        ob.Engine(ob.Config())
    with pytest.raises(ob.EngineError):
        ob.encode(b"alskdjas", True, None)


def test_engine_rejects_bad_cfg(engine_lib):
    from ofdm_b200 import engine
    c = engine.CCfg()
    engine_lib.ofdm_cfg_default(C.byref(c))
    c.nfft = 1024
    h = C.c_void_p()
    assert engine_lib.ofdm_engine_create(C.byref(c), 0, C.byref(h)) == -1
    assert b"nfft" in engine_lib.ofdm_last_error(None)


def test_header_and_wire_format():
    import ofdm_b200 as ob
    assert ob.Header(100).serialize() == (100).to_bytes(16, "little") and len(ob.Header(1).serialize()) == 16
    assert ob.Header.deserialize(ob.Header(12345).serialize()).packet_length == 12345
    sig = np.arange(10) + 1j * np.arange(10)
    b = ob.sig_to_bytes(sig)
    assert len(b) == 80
    np.testing.assert_array_equal(ob.bytes_to_sig(b + b"\x01\x02"), sig)           # chunks_exact drops the tail
    assert int(ob.ModulationScheme.Qam) == 2


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ofdm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                t = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in t.replace("the oracle has its own copy", ""), f"{f} mentions the oracle"


def test_shard_helpers():
    from ofdm_b200 import dist
    for n in (1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [dist.stream_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sh = dist.capture_shards(10_000_000, 8, 163840)
    assert sh[0].read_lo == 0 and sh[-1].read_hi == 10_000_000
    assert sh[0].own_lo == 0 and sh[-1].own_hi == 10_000_000 and all(a.own_hi == b.own_lo for a, b in zip(sh, sh[1:]))    # owned ranges partition
    assert all(a.read_hi - b.own_lo == 2 * 80 + 163840 and b.own_lo - b.read_lo == dist.CAPTURE_GUARD for a, b in zip(sh, sh[1:]))
    assert dist.err_rate([3, 2, 300, 0]) == 0.01


def test_rust_shim_declares_every_symbol():
    """rust/ofdm-sys/src/lib.rs (not compilable here: no rustc) must bind exactly the header's functions."""
    text = open(os.path.join(ROOT, "rust", "ofdm-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (ofdm_[a-z0-9_]+)\(", text)))
    assert rust == declared_functions()
    # struct field order of ofdm_cfg
    hdr = open(os.path.join(ROOT, "include", "ofdm_engine.h")).read()
    c_fields = re.findall(r"\b(?:uint32_t|const ofdm_fc32 \*)\s*(\w+);", hdr[hdr.index("typedef struct {\n    uint32_t struct_size"):hdr.index("} ofdm_cfg;")])
    r_fields = re.findall(r"pub (\w+):", text[text.index("pub struct ofdm_cfg"):text.index("pub struct ofdm_rx_diag")])
    assert c_fields == r_fields


def test_cpp_host_example_builds():
    from ofdm_b200 import _build
    _build.build_engine()
    exe = _build.build_host_example()
    assert os.access(exe, os.X_OK)


def test_header_is_plain_c99_and_links_from_c(tmp_path, engine_lib):
    """The boundary is a C ABI: include/ofdm_engine.h must compile as C99 (what cgo / bindgen / ctypesgen consume) and a C
    program must link against libofdm_b200.so and get sane answers from the calls that need no GPU."""
    import subprocess
    from ofdm_b200 import _build
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include "ofdm_engine.h"
#include <stdio.h>
int main(void) {
    ofdm_cfg c;
    ofdm_cfg_default(&c);
    c.modulation = OFDM_MOD_QAM64; c.guard_bands = 1; c.fec = 1;
    printf("%u %u %u %u %zu %zu %s\n", ofdm_abi_version(), ofdm_coded_len(&c, 576), ofdm_frame_data_syms(&c, 576), ofdm_frame_len(&c, 576),
           ofdm_rs_encoded_len(576), ofdm_rs_decoded_len(765), ofdm_status_name(OFDM_BAD_HEADER));
    return 0;
}
''')
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_build.LIB_PATH)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-o", str(exe),
                    "-L", libdir, "-lofdm_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[1:6] == ["1008", "29", "3120", "765", "892"] and out[6] == "BAD_HEADER"      # SURVEY.md 8d config 1 sizes


def test_tables_stdrng_known_answers(tmp_path):
    """The product's own StdRng restatement (ofdm_b200/csrc/tables.h, host C++) on the same known answers as the oracle's:
    the published ChaCha12 zero-key vector and rand 0.8's `test_stdrng_construction` values; and the two restatements give
    identical preamble / training tables (checked on the engine in tests/test_gpu_parity.py)."""
    import subprocess
    src = tmp_path / "kat.cpp"
    src.write_text(r'''
#include "tables.h"
#include <cstdio>
int main() {
    uint8_t zero[32] = {0};
    ofdm_host::StdRng z = ofdm_host::StdRng::from_seed(zero);
    for (int i = 0; i < 4; i++) { uint64_t v = z.next_u64(); for (int b = 0; b < 8; b++) printf("%02x", (unsigned)((v >> (8 * b)) & 255)); }
    printf("\n");
    uint8_t seed[32] = {1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0};
    ofdm_host::StdRng g = ofdm_host::StdRng::from_seed(seed);
    printf("%llu\n", (unsigned long long)g.next_u64());
    uint8_t seed1[32];
    for (int i = 0; i < 4; i++) { uint64_t v = g.next_u64(); for (int b = 0; b < 8; b++) seed1[8 * i + b] = (uint8_t)(v >> (8 * b)); }
    ofdm_host::StdRng g1 = ofdm_host::StdRng::from_seed(seed1);
    printf("%llu\n", (unsigned long long)g1.next_u64());
    return 0;
}
''')
    exe = tmp_path / "kat"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "ofdm_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
    assert out[1:] == ["10719222850664546238", "14064965282130556830"]


def test_rust_build_script_compiles_every_translation_unit():
    """rust/ofdm-sys/build.rs (cannot be run here: no cargo) must list exactly the units ofdm_b200/_build.py links, with the
    array length Rust's type needs -- a unit missing there is an undefined symbol at the maintainer's first `cargo build`."""
    import re
    from ofdm_b200 import _build
    src = open(os.path.join(ROOT, "rust", "ofdm-sys", "build.rs")).read()
    m = re.search(r"const UNITS: \[&str; (\d+)\] = \[(.*?)\];", src, re.S)
    units = re.findall(r'"(\w+)"', m.group(2))
    assert int(m.group(1)) == len(units)
    assert sorted(units) == sorted(_build.UNITS)
    for u in units:
        assert os.path.exists(os.path.join(ROOT, "ofdm_b200", "csrc", u + ".cu"))
