"""Payload adapters of the reference's lab demos: text corpus and palette image.

create_transmission_text / decipher_transmission_text (src/utils.rs:87-95,139-150) and
decipher_transmision_colorspace (src/utils.rs:182-205) wrap the payload bytes of the modem; with `ecc` they go through the
RS(255,223) outer code, which runs on the GPU (ofdm_rs_encode_batch / ofdm_rs_decode_batch). The byte <-> colour table is
the xterm 256-colour palette (src/packets/colors.rs:7-9 loads it from support/colors.json); it is generated here from
xterm's defining rule (16 system colours, a 6x6x6 cube with levels 0,95,135,175,215,255, 24 greys 8+10i), which
reproduces the reference file's 256 (id, r, g, b) entries exactly.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import api

# Percy Bysshe Shelley, "Ozymandias" (1818, public domain), with the line breaks and punctuation of the reference's
# CORPUS constant (src/utils.rs:70-85): leading and trailing newline, typographic dash / opening quote, spaced ellipsis.
CORPUS = (
    "\nI met a traveller from an antique land,\n"
    "Who said—“Two vast and trunkless legs of stone\n"
    "Stand in the desert. . . . Near them, on the sand,\n"
    "Half sunk a shattered visage lies, whose frown,\n"
    "And wrinkled lip, and sneer of cold command,\n"
    "Tell that its sculptor well those passions read\n"
    "Which yet survive, stamped on these lifeless things,\n"
    "The hand that mocked them, and the heart that fed;\n"
    "And on the pedestal, these words appear:\n"
    "My name is Ozymandias, King of Kings;\n"
    "Look on my Works, ye Mighty, and despair!\n"
    "Nothing beside remains. Round the decay\n"
    "Of that colossal Wreck, boundless and bare\n"
    "The lone and level sands stretch far away.\n"
)

_SYSTEM = [(0, 0, 0), (128, 0, 0), (0, 128, 0), (128, 128, 0), (0, 0, 128), (128, 0, 128), (0, 128, 128), (192, 192, 192),
           (128, 128, 128), (255, 0, 0), (0, 255, 0), (255, 255, 0), (0, 0, 255), (255, 0, 255), (0, 255, 255), (255, 255, 255)]
_LEVELS = (0, 95, 135, 175, 215, 255)


def xterm256_palette() -> np.ndarray:
    """uint8 [256, 3]: colour id -> (r, g, b)."""
    pal = list(_SYSTEM)
    pal += [(r, g, b) for r in _LEVELS for g in _LEVELS for b in _LEVELS]
    pal += [(8 + 10 * i,) * 3 for i in range(24)]
    return np.array(pal, np.uint8)


_PALETTE = xterm256_palette()


def create_transmission_text(msg_bytes: int, ecc: bool, corpus: str = CORPUS, device: int = 0) -> bytes:
    """src/utils.rs:87-95: the corpus bytes cycled to msg_bytes, RS-coded when `ecc`."""
    raw = corpus.encode("utf-8")
    body = (raw * (msg_bytes // len(raw) + 1))[:msg_bytes]
    return api.create_transmission_bytes(body, device) if ecc else body


def decipher_transmission_text(num_bytes: int, data: bytes, ecc: bool, device: int = 0) -> Optional[str]:
    """src/utils.rs:139-150: None when the outer code fails or the bytes are not UTF-8."""
    if ecc:
        out = api.decipher_transmission_bytes(data, device)
        if out is None:
            return None
        data = out[:num_bytes]
    try:
        return bytes(data).decode("utf-8")
    except UnicodeDecodeError:
        return None


def decipher_transmision_colorspace(data: bytes, ecc: bool, device: int = 0) -> Optional[np.ndarray]:
    """src/utils.rs:182-205: one 0x00RRGGBB word per received byte (after the outer code when `ecc`)."""
    if ecc:
        data = api.decipher_transmission_bytes(data, device)
        if data is None:
            return None
    ids = np.frombuffer(bytes(data), np.uint8)
    rgb = _PALETTE[ids].astype(np.uint32)
    return (rgb[:, 0] << 16) | (rgb[:, 1] << 8) | rgb[:, 2]


def closest_color_ids(rgb: np.ndarray) -> np.ndarray:
    """ColorMap::get_closest (src/packets/colors.rs:40-43) for an array of (r, g, b): nearest palette entry in RGB
    distance; among equidistant entries (the palette repeats e.g. black as 0 and 16) the lowest id."""
    px = np.asarray(rgb, np.int32).reshape(-1, 3)
    d = ((px[:, None, :] - _PALETTE[None, :, :].astype(np.int32)) ** 2).sum(axis=2)
    return d.argmin(axis=1).astype(np.uint8)
