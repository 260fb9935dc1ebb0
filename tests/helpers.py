"""Shared helpers of the parity tests: seeded action streams (the reference's two input distributions)."""
import numpy as np

SIZES = [3, 3, 3, 2, 2, 3]


def encode(heads):
    """MegaverseGym::setActions (megaverse.cpp:100-116)"""
    mask, idx = 0, 0
    for a, sz in zip(heads, SIZES):
        if a > 0:
            mask |= 1 << (idx + int(a))
        idx += sz - 1
    return mask


def random_bit_actions(rng, n):
    """megaverse_test_app.cpp:140-147: one uniformly random action bit per agent (bit 0 = no-op)"""
    return (1 << rng.integers(0, 11, size=n)).astype(np.int32)


def random_head_actions(rng, n):
    """action_space.sample(): independent uniform sample of each Discrete head (megaverse/tests/test_env.py:12-13)"""
    out = np.zeros(n, dtype=np.int32)
    for i in range(n):
        out[i] = encode([rng.integers(0, s) for s in SIZES])
    return out


def purposeful_actions(rng, n, step):
    """mostly walk forward / turn, interact often: makes pick-ups, placements and rewards actually happen"""
    out = np.zeros(n, dtype=np.int32)
    for i in range(n):
        r = rng.random()
        m = 0
        if r < 0.7:
            m |= 1 << 3  # forward
        if rng.random() < 0.25:
            m |= (1 << 5) if rng.random() < 0.5 else (1 << 6)
        if rng.random() < 0.3:
            m |= 1 << 8  # interact
        if rng.random() < 0.1:
            m |= 1 << 7  # jump
        if rng.random() < 0.1:
            m |= (1 << 9) if rng.random() < 0.6 else (1 << 10)
        out[i] = m
    return out


def rearrange_controller(o, e, A):
    """one action mask per agent of env e: fetch a misplaced object, carry it to its target cell on the work pedestal"""
    st = o.state(e); arr = o.arrangement(e)
    n = arr[0]; items = arr[1:1+5*n].reshape(n,5); m = arr[1+5*n]; objs = arr[2+5*n:2+5*n+2*m].reshape(m,2); wc = arr[2+5*n+2*m:]
    nobj = int(st[5]); base = 8 + 26*A
    O = st[base:base+9*nobj].reshape(nobj,9)
    out = np.zeros(A, dtype=np.int32)
    # which target offsets are already satisfied / free
    vox = np.floor(O[:, :3]).astype(int) - wc
    for a in range(A):
        ag = st[8+26*a: 8+26*(a+1)]
        pos = ag[0:3]; b = ag[3:12]; fwd = np.array([b[6], -b[8]]); fwd /= (np.linalg.norm(fwd)+1e-9)
        carrying = int(ag[22])
        if a > 0:  # the other agents wander
            out[a] = 1 << 3 if (int(st[2]) // 7 + a) % 3 else 1 << 5
            continue
        if carrying < 0:
            # choose an object that is not at a matching offset
            cand = None
            for i in range(nobj):
                ok = any((items[q,0]==objs[i,0] and items[q,1]==objs[i,1] and (items[q,2:5]==vox[i]).all()) for q in range(n))
                if not ok and O[i,6] < 0:
                    cand = i; break
            if cand is None:
                continue
            tgt = O[cand, [0,2]]; want_dist = 1.0
        else:
            i = carrying
            # free target offset of the same shape+colour
            tgt = None
            for q in range(n):
                if items[q,0]==objs[i,0] and items[q,1]==objs[i,1]:
                    taken = any((vox[j]==items[q,2:5]).all() and O[j,6] < 0 for j in range(nobj))
                    if not taken and items[q,3]==0:
                        tgt = (items[q,2:5]+wc)[[0,2]] + 0.5; break
            if tgt is None:
                continue
            want_dist = 1.0
        d = tgt - pos[[0,2]]; dist = np.linalg.norm(d)
        ang = np.arctan2(fwd[0]*d[1]-fwd[1]*d[0], fwd[0]*d[0]+fwd[1]*d[1])
        mask = 0
        if abs(ang) > 0.12:
            mask |= (1<<6) if ang > 0 else (1<<5)
        elif dist > want_dist + 0.15:
            mask |= 1<<3
        elif dist < want_dist - 0.25:
            mask |= 1<<4
        else:
            mask |= 1<<8
        out[a] = mask
    return out

