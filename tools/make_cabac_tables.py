"""Extracts the H.264 CABAC constant tables (ITU-T H.264 Tables 9-12..9-33 context initialisation for I slices,
Table 9-44 rangeTabLPS, Table 9-45 transIdxLPS/MPS, Table 9-43 8x8 ctxIdxInc maps) as plain numbers from the
reference's Rust source, where they appear as literal arrays (src/video/cabac/table.rs:4-1172,
src/video/cabac/consts.rs:135-213, :402-466), into dryv_b200/csrc/cabac_tables.json. They are standard constants, not
code; the stream writer in tests/avc/ (test tooling for the libavcodec cross-check of the oracle) loads the JSON.

    python tools/make_cabac_tables.py        # needs /root/reference, run in the build container only
"""
import json
import os
import re

REF = "/root/reference/src/video/cabac"
HERE = os.path.dirname(os.path.abspath(__file__))


def array_body(text, name):
    i = text.index(name)
    i = text.index("=", i)
    depth, j = 0, i
    while True:
        c = text[j]
        if c == "[":
            depth += 1
        elif c == "]":
            depth -= 1
            if depth == 0:
                break
        j += 1
    body = re.sub(r"/\*.*?\*/", "", text[i:j + 1], flags=re.S)
    return re.sub(r"//[^\n]*", "", body)


def ints(s):
    return [int(v) for v in re.findall(r"-?\d+", s)]


table = open(os.path.join(REF, "table.rs")).read()
consts = open(os.path.join(REF, "consts.rs")).read()
rows = re.findall(r"\[\(\s*(-?\d+),\s*(-?\d+)\)", array_body(table, "CTX_INIT_TABLE"))
ctx_init_i = [[int(m), int(n)] for m, n in rows]
assert len(ctx_init_i) == 1031, len(ctx_init_i)
rng = ints(array_body(consts, "RANGE_TAB_LPS"))
assert len(rng) == 256
lps, mps = ints(array_body(consts, "TRANS_IDX_LPS")), ints(array_body(consts, "TRANS_IDX_MPS"))
assert len(lps) == 64 and len(mps) == 64
tab8 = ints(array_body(consts, "SIGNIFICANT_COEFF_FLAG_TAB8X8"))
assert len(tab8) == 63 * 3
out = {
    "source": "ITU-T H.264 Tables 9-12..9-33 (I-slice column), 9-43, 9-44, 9-45",
    "ctx_init_i": ctx_init_i,
    "range_tab_lps": [rng[4 * i:4 * i + 4] for i in range(64)],
    "trans_idx_lps": lps,
    "trans_idx_mps": mps,
    "sig8x8_frame": tab8[0::3],
    "last8x8": tab8[2::3],
}
with open(os.path.join(HERE, "..", "dryv_b200", "csrc", "cabac_tables.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("ctx", len(ctx_init_i), "first", ctx_init_i[:4], "range[0]", out["range_tab_lps"][0], "sig8", out["sig8x8_frame"][:8], "last8", out["last8x8"][:8])
