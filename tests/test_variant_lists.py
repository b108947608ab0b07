"""The shape ids tools/r2_first_call.sh and tests/test_experimental_gpu.py ask for must be compiled in (msgpu.cu falls back to the
default shape for an unknown id - a typo would silently measure the default twice)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compiled():
    src = open(os.path.join(ROOT, "libmspack_b200", "csrc", "msgpu.cu")).read()
    lzx_block = src[src.index("#define LZXC_VARIANTS(X)"):src.index("#define ZIPK_NT")]
    lzx = {int(m) for m in re.findall(r"X\((\d+),", lzx_block)}
    zip_block = src[src.index("#define ZIPC_VARIANTS(X)"):src.index("#define ZIP_VARIANT_OPT1")]
    zips = {int(m) for m in re.findall(r"X\((\d+),", zip_block)}
    base = int(re.search(r"#define ZIP_VARIANT_OPT1 (\d+)", src).group(1))
    top = max(int(m) for m in re.findall(r"zip_variant <= ZIP_VARIANT_OPT1 \+ (\d+)", src))
    zips |= set(range(base, base + top + 1))
    qtm = {0} | {int(m) for m in re.findall(r"qtm_variant == (\d+)\) k_p1_qtm", src)}
    return lzx, zips, qtm


def test_requested_shapes_exist():
    lzx, zips, qtm = _compiled()
    sh = open(os.path.join(ROOT, "tools", "r2_first_call.sh")).read()
    for line in sh.splitlines():
        m = re.search(r"variant_bench\.py (\d+) ([\d ]+?) >", line)
        if not m:
            continue
        ids = {int(x) for x in m.group(2).split()}
        have = zips if "VB_CODEC=1" in line else (qtm if "VB_CODEC=2" in line else lzx)
        assert ids <= have, (line[:80], sorted(ids - have))
    import test_experimental_gpu as T
    for env, v, _ in T.SHAPES:
        have = {"MSGPU_LZX_VARIANT": lzx, "MSGPU_ZIP_VARIANT": zips, "MSGPU_QTM_VARIANT": qtm, "MSGPU_P2_VARIANT": {0, 1}}[env]
        assert v in have, (env, v)
