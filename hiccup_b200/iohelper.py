"""Bit-string <-> padded bytes, the `.hic` bit payload framing.

Same functions and results as reference hiccup/iohelper.py:35-56, without the `bitstring`
dependency: byte 0 is the pad count p = 8 - (n mod 8) (1..8 -- an aligned string gets a whole zero
byte), then the bits MSB-first, then p zero bits.
"""


def padded_bs_2_bytes(s: str) -> bytes:
    pad = 8 - (len(s) % 8)
    framed = format(pad, "08b") + s + "0" * pad
    return int(framed, 2).to_bytes(len(framed) // 8, "big")


def padded_bytes_2_bs(bites) -> str:
    data = bytes(bites)
    n_bits = payload_bit_count(data)
    body = data[1:]
    if not body:
        return ""
    return bin(int.from_bytes(body, "big"))[2:].zfill(8 * len(body))[:n_bits]


def payload_bit_count(data) -> int:
    """Number of payload bits in a framed byte string (what padded_bytes_2_bs would return the
    length of), following the reference's arithmetic: the pad byte is read as a signed int8."""
    pad = data[0] - 256 if data[0] >= 128 else data[0]
    return max(0, 8 * (len(data) - 1) - pad)
