"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on CPU.

The oracle (oracle/) is the CHECKER: tests may import it, the product (libzkp_b200/) may not.
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "groth16_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def co():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def po():
    from oracle import zkp_oracle
    return zkp_oracle


@pytest.fixture(scope="session")
def trapdoor(po):
    td = po.Trapdoor.from_seed(1)
    return (td.alpha, td.beta, td.gamma, td.delta, td.tau)


class Keys:
    def __init__(self, circuit, pk_bytes, vk_bytes, opk):
        self.circuit, self.pk_bytes, self.vk_bytes, self.opk = circuit, pk_bytes, vk_bytes, opk


@pytest.fixture(scope="session")
def eq_keys(co, trapdoor):
    c = co.Circuit("equality")
    pk, vk = c.setup(trapdoor)
    return Keys(c, pk, vk, co.ProvingKey(pk))


@pytest.fixture(scope="session")
def mb_keys(co, trapdoor):
    c = co.Circuit("membership")
    pk, vk = c.setup(trapdoor)
    return Keys(c, pk, vk, co.ProvingKey(pk))


def fr_rand(po, seed, n):
    rng = po.SplitMix64(seed)
    return np.frombuffer(b"".join(rng.next_fr().to_bytes(32, "little") for _ in range(n)), np.uint8).reshape(n, 32).copy()


@pytest.fixture(scope="session")
def frs(po):
    return lambda seed, n: fr_rand(po, seed, n)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
