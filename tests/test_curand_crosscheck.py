"""An INDEPENDENT pin of the Philox4x32-10 generator: NVIDIA's own implementation in cuRAND's
header `curand_philox4x32_x.h` (curand_Philox4x32_10), compiled for the host from the CUDA toolkit's include directory,
against the oracle's generator (CPU) and against the device generator of the kernel library (GPU).  Nothing of cuRAND is
linked or copied: the header is compiled where it lies, by a 15-line driver written here, and only its outputs are used.
Together with the three Random123 known-answer vectors (tests/golden/philox_kats.json) this fixes the draw source of the
whole path by two parties that are not the builder."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

DRIVER = r"""
#include <cstdio>
#include <vector_types.h>
#define QUALIFIERS static inline
#include <curand_philox4x32_x.h>
int main() {
    unsigned v[6];
    while (scanf("%u %u %u %u %u %u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6) {
        uint4 c; c.x = v[0]; c.y = v[1]; c.z = v[2]; c.w = v[3];
        uint2 k; k.x = v[4]; k.y = v[5];
        const uint4 r = curand_Philox4x32_10(c, k);
        printf("%u %u %u %u\n", r.x, r.y, r.z, r.w);
    }
    return 0;
}
"""


def cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda", "/usr/local/cuda-12.9"):
        if d and os.path.exists(os.path.join(d, "include", "curand_philox4x32_x.h")):
            return os.path.join(d, "include")
    nvcc = shutil.which("nvcc")
    if nvcc:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
        if os.path.exists(os.path.join(d, "curand_philox4x32_x.h")):
            return d
    return None


def vectors():
    rng = np.random.default_rng(20261020)
    v = rng.integers(0, 2 ** 32, size=(3000, 6), dtype=np.uint64)
    edge = [[0] * 6, [0xFFFFFFFF] * 6, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0],
            [1, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 0, (2 << 28) | 5, 7, 1729, 0]]
    return np.concatenate([np.array(edge, dtype=np.uint64), v])


@pytest.fixture(scope="module")
def curand_blocks(tmp_path_factory):
    inc, gxx = cuda_include(), shutil.which("g++")
    if not inc or not gxx:
        pytest.skip("needs g++ and the CUDA toolkit headers")
    d = tmp_path_factory.mktemp("curand")
    (d / "driver.cpp").write_text(DRIVER)
    subprocess.run([gxx, "-O1", f"-I{inc}", "-o", str(d / "driver"), str(d / "driver.cpp")], check=True, capture_output=True)
    v = vectors()
    text = "\n".join(" ".join(str(int(x)) for x in row) for row in v) + "\n"
    r = subprocess.run([str(d / "driver")], input=text, capture_output=True, text=True, check=True)
    out = np.array([[int(x) for x in ln.split()] for ln in r.stdout.splitlines()], dtype=np.uint64)
    assert out.shape == (v.shape[0], 4)
    return v, out


def test_curand_reproduces_the_random123_vectors(curand_blocks):
    v, out = curand_blocks
    assert [hex(int(x)) for x in out[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(int(x)) for x in out[1]] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(int(x)) for x in out[2]] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_oracle_generator_equals_curand(curand_blocks):
    v, out = curand_blocks
    for row, want in zip(v, out):
        got = O.philox4x32_10([int(x) for x in row[:4]], [int(x) for x in row[4:]])
        assert got == [int(x) for x in want], row


@pytest.mark.gpu
def test_device_generator_equals_curand(ctx, curand_blocks):
    v, out = curand_blocks
    keys = {(int(r[4]), int(r[5])) for r in v}
    for key in list(keys)[:64]:                       # eqs_philox_blocks takes one key per call
        sel = np.flatnonzero((v[:, 4] == key[0]) & (v[:, 5] == key[1]))
        got = ctx.philox_blocks(v[sel, :4].astype(np.uint32), np.array(key, dtype=np.uint32))
        assert np.array_equal(got.astype(np.uint64), out[sel])
    # many counters under one key (the shape the kernels use): (index lo, index hi, stream << 28 | j, pass)
    rng = np.random.default_rng(3)
    ctr = rng.integers(0, 2 ** 32, size=(4096, 4), dtype=np.uint64)
    key = (1729, 0)
    got = ctx.philox_blocks(ctr.astype(np.uint32), np.array(key, dtype=np.uint32))
    for i in range(0, 4096, 97):
        assert O.philox4x32_10([int(x) for x in ctr[i]], list(key)) == [int(x) for x in got[i]]
