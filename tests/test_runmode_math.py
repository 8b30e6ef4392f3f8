"""The Golomb-Rice kernels decide which samples are coded in run mode (ffv1enc.c:327-357) with a carry chain evaluated by
one 64-bit addition per 32 samples (gr_run_members in csrc/ffv1_ctx_replay.cu, k_gr_pack in csrc/ffv1_enc_kernels.cu).
This checks that formula against the sequential definition of the reference's encode_line on random lines."""
import numpy as np

def members_sequential(ctx0, zero):
    """M(i) = run_mode after `if (context == 0) run_mode = 1` when sample i is coded"""
    out, run_mode = [], 0
    for c, z in zip(ctx0, zero):
        if c:
            run_mode = 1
        out.append(run_mode)
        if run_mode and not z:          # a non-zero residual ends the run
            run_mode = 0
    return out

def members_carry_chain(ctx0, zero):
    out, carry = [], 0
    for g in range(0, len(ctx0), 32):
        c = sum(int(b) << i for i, b in enumerate(ctx0[g:g + 32]))
        z = sum(int(b) << i for i, b in enumerate(zero[g:g + 32]))
        a, b = z, z & c
        s = a + b + carry
        chain = (s & 0xFFFFFFFF) ^ a ^ b
        carry = s >> 32
        m = c | chain
        out += [(m >> i) & 1 for i in range(len(ctx0[g:g + 32]))]
    return out

def test_run_membership_carry_chain_equals_sequential():
    rng = np.random.default_rng(7)
    for trial in range(300):
        n = 32 * int(rng.integers(1, 12))                      # whole groups: a line's tail lanes carry zero = ctx0 = 0
        p0, pz = rng.choice([0.02, 0.1, 0.5]), rng.choice([0.1, 0.5, 0.95])
        ctx0 = (rng.random(n) < p0).astype(int).tolist()
        zero = (rng.random(n) < pz).astype(int).tolist()
        assert members_carry_chain(ctx0, zero) == members_sequential(ctx0, zero), trial

def test_zero_length_run_terminators_walk_run_index_down():
    """k_gr_pack's parallel path: in a group without absorbed samples the k-th run-mode sample sees
    run_index = max(r0 - k, 0) and emits 1 + log2_run[run_index] zero bits (ffv1enc.c:338-346)"""
    log2_run = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7] + list(range(8, 25))
    rng = np.random.default_rng(8)
    for trial in range(200):
        r0 = int(rng.integers(0, 12))
        inrun = (rng.random(32) < 0.2).astype(int)
        run_index, bits_seq = r0, []
        for m in inrun:                                        # sequential reference: run_count is 0, residual non-zero
            if m:
                bits_seq.append(1 + log2_run[run_index])
                if run_index:
                    run_index -= 1
            else:
                bits_seq.append(0)
        before = np.concatenate([[0], np.cumsum(inrun)[:-1]])
        bits_par = [(1 + log2_run[max(r0 - int(b), 0)]) if m else 0 for m, b in zip(inrun, before)]
        assert bits_par == bits_seq and run_index == max(r0 - int(inrun.sum()), 0)
