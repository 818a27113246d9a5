"""Extracts a few G1 points from the KZG SRS the REFERENCE ships (params/kzg_bn254_8.srs: halo2 ParamsKZG raw format --
u32 k, then 2^k G1 points as 64 raw bytes (x | y, little-endian Montgomery-form Fq limbs), ...) into a small fixture.
They are [tau^i]G on BN254, so their group order is the scalar-field modulus r: an independent, reference-shipped pin
of the Fr modulus our whole path computes in (tests/test_oracle.py::test_fr_modulus_is_the_order_of_the_reference_srs_points).

    python tests/golden/make_srs_points.py        (run where /root/reference exists)
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRS = "/root/reference/params/kzg_bn254_8.srs"

if __name__ == "__main__":
    d = open(SRS, "rb").read()
    k = int.from_bytes(d[:4], "little")
    assert len(d) == 4 + 2 * (1 << k) * 64 + 2 * 128, "unexpected SRS layout"
    pts = []
    for i in (0, 1, 2, 3, 255):
        b = d[4 + 64 * i: 4 + 64 * (i + 1)]
        pts.append({"index": i, "x_mont_le_hex": b[:32].hex(), "y_mont_le_hex": b[32:].hex()})
    json.dump({"source": "params/kzg_bn254_8.srs (reference), g[i] = [tau^i]G1, raw Montgomery-form Fq coordinates",
               "k": k, "points": pts}, open(os.path.join(HERE, "srs_g1_points.json"), "w"), indent=1)
    print("wrote srs_g1_points.json")
