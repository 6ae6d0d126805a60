"""CPU oracle for the go-curdleproofs G1 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import, link or execute it, and only as the
checker (never as the thing measured as the GPU path, never shipped).

PARITY STATUS: **unpinned against a Go run.**  The reference
(/root/reference, Go) cannot be compiled here (no Go toolchain, and its
arithmetic lives in the un-vendored dependency
github.com/consensys/gnark-crypto v0.11.0, go.mod:6; transcript in
github.com/jsign/merlin v0.0.0-20230603163309-c45ec8d8b2ce, go.mod:7; SHAKE256
from golang.org/x/crypto, go.mod:9).  The reference's own tests contain no
golden vectors.  The oracle is therefore pinned to
  * public known-answer tests of every primitive it restates (Merlin's
    "test protocol" vector, SHA3/SHAKE via hashlib, the BLS12-381 generator
    encodings 97f1d3a7…c6bb / a572cbea…0f4e, curve/subgroup identities), and
  * two independent restatements agreeing byte for byte (pure-Python big-int
    in ``oracle/*.py`` and plain C 6x64-limb Montgomery in ``oracle/c/``).
Three assumptions only a real Go run could close are listed in DESIGN.md
("assumed, unpinned"): gnark's uint32 big-endian slice-length prefix, decoder
strictness on non-canonical inputs, and Go math/rand permutations used by the
reference's benchmark ``setup`` (we substitute common.Rand.GeneratePermutation).
"""
