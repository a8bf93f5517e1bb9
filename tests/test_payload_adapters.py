"""Payload adapters: text corpus and palette image around the modem payload."""
import numpy as np
import pytest


def test_xterm_palette_rule():
    from ofdm_b200.payloads import xterm256_palette, closest_color_ids, decipher_transmision_colorspace
    pal = xterm256_palette()
    assert pal.shape == (256, 3)
    assert tuple(pal[0]) == (0, 0, 0) and tuple(pal[9]) == (255, 0, 0) and tuple(pal[15]) == (255, 255, 255)
    assert tuple(pal[16]) == (0, 0, 0) and tuple(pal[21]) == (0, 0, 255) and tuple(pal[196]) == (255, 0, 0) and tuple(pal[231]) == (255, 255, 255)
    assert tuple(pal[232]) == (8, 8, 8) and tuple(pal[255]) == (238, 238, 238)
    words = decipher_transmision_colorspace(bytes([0, 9, 21, 255]), ecc=False)
    assert [hex(int(w)) for w in words] == ["0x0", "0xff0000", "0xff", "0xeeeeee"]
    ids = closest_color_ids(pal)                                         # every palette colour maps to its first occurrence
    for i, k in enumerate(ids):
        assert tuple(pal[k]) == tuple(pal[i]) and k <= i
    assert int(closest_color_ids([[250, 5, 3]])[0]) == 9


def test_text_corpus_cycles():
    from ofdm_b200.payloads import create_transmission_text, decipher_transmission_text, CORPUS
    raw = CORPUS.encode()
    assert raw.startswith(b"\nI met a traveller") and raw.endswith(b"stretch far away.\n")
    body = create_transmission_text(1024, ecc=False)
    assert len(body) == 1024 and body[:len(raw)] == raw and body[len(raw):2 * len(raw)] == raw[:1024 - len(raw)]
    txt = decipher_transmission_text(1024, body[:len(raw)], ecc=False)
    assert txt == CORPUS
    assert decipher_transmission_text(4, b"\xff\xfe\x00\x01", ecc=False) is None      # String::from_utf8(..).ok()


@pytest.mark.gpu
def test_ecc_packets_round_trip():
    # the reference's `ecc_packets` test (src/utils.rs:358-367) with assertions instead of dbg!
    from ofdm_b200.payloads import create_transmission_text, decipher_transmission_text, decipher_transmision_colorspace, CORPUS
    ecced = create_transmission_text(1024, True)
    assert len(ecced) == 255 * (1024 // 223 + 1)
    bad = bytearray(ecced)
    for b in range(len(bad) // 255):
        for k in range(0, 255, 17):                                      # 15 symbol errors per block
            bad[255 * b + k] ^= 0xA5
    text = decipher_transmission_text(1024, bytes(bad), True)
    raw = CORPUS.encode()
    assert text is not None and text.encode() == (raw * 3)[:1024]
    words = decipher_transmision_colorspace(bytes(bad), True)
    assert words is not None and words.size == 223 * (len(bad) // 255 + 1)
    bad[1] ^= 1; bad[2] ^= 1
    assert decipher_transmission_text(1024, bytes(bad), True) is None    # 17 errors in block 0
