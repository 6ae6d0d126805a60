// Cooperative fibers for the host-side Fiat-Shamir work of a batch.
//
// The per-proof host code between two GPU stages (Merlin / STROBE appends and challenge draws with
// rejection sampling, common.Rand's SHAKE256, Fr folds) is a strictly serial sponge per proof, and
// rejection sampling makes the proofs of a batch drift apart, so the transcripts cannot simply run
// in lock step.  Instead up to eight proofs run as fibers on one pool thread: a fiber that needs a
// Keccak-f[1600] permutation parks its state and yields; once every live fiber of the group is
// parked the scheduler permutes all parked states with ONE eight-way AVX-512 call
// (host/keccak_x8.cpp) and resumes them.  The per-proof code is unchanged - the hook sits inside
// keccak_f1600() - and every proof still sees exactly its own sequence of permutations.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace cdlh {

// true when eight-way hashing is usable on this host (AVX-512F) and not disabled (CDL_NO_FIBERS)
bool fibers_available();
// Called by keccak_f1600(): when running inside a fiber, parks `st` until the group's next
// eight-way permutation and returns true; returns false outside a fiber (caller permutes itself).
bool fiber_keccak(uint64_t* st);
// fn(i) for i in [first, first + count), count <= 8, as cooperating fibers on the calling thread
void run_fiber_group(const std::function<void(size_t)>& fn, size_t first, size_t count);

}  // namespace cdlh
