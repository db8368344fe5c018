"""Address patterns for tools/l1_gather_probe.cu (stdout -> its stdin).

A pattern gives every lane a (line, slot): byte offset 128*line + vec*slot inside the CTA's window.  The sets below
isolate what a warp-wide gather pays for: distinct entries per quarter-warp, same / different 128-byte lines, same /
different slots (bank groups), how lanes are grouped, sharing between quarter-warps.
"""
import sys


def emit(name, vec, lanes):
    assert len(lanes) == 32
    print(name, vec, " ".join(f"{l}:{s}" for l, s in lanes))


def per_quarter(fn):
    """fn(q, j) -> (line, slot) for lane j of quarter-warp q."""
    return [fn(l // 8, l % 8) for l in range(32)]


def main():
    for vec in (32, 16):
        S = 128 // vec
        tag = f"v{vec}"
        print(f"# ---- {vec}-byte loads, {S} slots per 128-byte line")
        emit(f"{tag}_bcast_all", vec, [(0, 0)] * 32)
        emit(f"{tag}_coalesced", vec, [(l // S, l % S) for l in range(32)])
        # one entry per quarter
        emit(f"{tag}_q1_sameslot_difflines", vec, per_quarter(lambda q, j: (8 * q, 0)))
        emit(f"{tag}_q1_diffslot_difflines", vec, per_quarter(lambda q, j: (8 * q, q % S)))
        emit(f"{tag}_q1_sameline", vec, per_quarter(lambda q, j: (0, q % S)))
        # k distinct entries per quarter, blocked lanes (j*k//8) or interleaved lanes (j % k)
        for k in (2, 3, 4, 5, 6, 8):
            for group, gname in ((lambda j, k=k: j * k // 8, "blk"), (lambda j, k=k: j % k, "ilv")):
                emit(f"{tag}_q{k}{gname}_sameline", vec, per_quarter(lambda q, j: (8 * q + group(j) // S, group(j) % S)))
                emit(f"{tag}_q{k}{gname}_difflines_diffslots", vec, per_quarter(lambda q, j: (8 * q + group(j), group(j) % S)))
                emit(f"{tag}_q{k}{gname}_difflines_sameslot", vec, per_quarter(lambda q, j: (8 * q + group(j), 0)))
        # 4 distinct with slot multiset variations (different lines)
        for slots in ((0, 0, 1, 1), (0, 1, 2, 0), (0, 0, 0, 1), (0, 2, 0, 2), (0, 1, 0, 1)):
            emit(f"{tag}_q4blk_slots" + "".join(map(str, slots)), vec,
                 per_quarter(lambda q, j: (8 * q + j // 2, slots[j // 2] % S)))
        # sharing between quarters: every quarter reads the same 4 entries (4 lines, 4 slots)
        emit(f"{tag}_q4blk_shared_between_quarters", vec, per_quarter(lambda q, j: (j // 2, (j // 2) % S)))
        emit(f"{tag}_q4blk_shared_sameslot", vec, per_quarter(lambda q, j: (j // 2, 0)))
        # half-warp structure: 16 lanes = 4 distinct entries in lane blocks of 4
        emit(f"{tag}_h4blk_difflines_diffslots", vec, [((l // 16) * 8 + (l % 16) // 4, ((l % 16) // 4) % S) for l in range(32)])
        emit(f"{tag}_h8blk_difflines_diffslots", vec, [((l // 16) * 8 + (l % 16) // 2, ((l % 16) // 2) % S) for l in range(32)])
        # whole warp: n distinct entries in lane blocks, different lines, slots cycling
        for n in (2, 4, 8, 16):
            emit(f"{tag}_w{n}blk_difflines_diffslots", vec, [(l * n // 32, (l * n // 32) % S) for l in range(32)])
            emit(f"{tag}_w{n}blk_difflines_sameslot", vec, [(l * n // 32, 0) for l in range(32)])
            emit(f"{tag}_w{n}blk_sameline" if n <= S else f"{tag}_w{n}blk_packedlines", vec,
                 [((l * n // 32) // S, (l * n // 32) % S) for l in range(32)])
        # odd offsets inside a line: two entries of one line that straddle halves
        if vec == 16:
            emit(f"{tag}_q2blk_sameline_slots0_1", vec, per_quarter(lambda q, j: (8 * q, j // 4)))
            emit(f"{tag}_q2blk_sameline_slots0_2", vec, per_quarter(lambda q, j: (8 * q, 2 * (j // 4))))
            emit(f"{tag}_q2blk_difflines_slots0_1", vec, per_quarter(lambda q, j: (8 * q + j // 4, j // 4)))
            emit(f"{tag}_q2blk_difflines_slots0_2", vec, per_quarter(lambda q, j: (8 * q + j // 4, 2 * (j // 4))))
            emit(f"{tag}_q2blk_difflines_slots0_4", vec, per_quarter(lambda q, j: (8 * q + j // 4, 4 * (j // 4))))
    for vec in (8, 4):
        S = 128 // vec
        tag = f"v{vec}"
        print(f"# ---- {vec}-byte loads")
        emit(f"{tag}_bcast_all", vec, [(0, 0)] * 32)
        emit(f"{tag}_coalesced", vec, [(l // S, l % S) for l in range(32)])
        emit(f"{tag}_w8blk_difflines_diffslots", vec, [(l // 4, (l // 4) * (S // 8)) for l in range(32)])
        emit(f"{tag}_w8blk_difflines_sameslot", vec, [(l // 4, 0) for l in range(32)])
        emit(f"{tag}_w16blk_difflines_diffslots", vec, [(l // 2, (l // 2) * (S // 16)) for l in range(32)])
        emit(f"{tag}_w16blk_difflines_sameslot", vec, [(l // 2, 0) for l in range(32)])
        emit(f"{tag}_w32_difflines_diffslots", vec, [(l, l * (S // 32) if S >= 32 else l % S) for l in range(32)])
        emit(f"{tag}_w32_difflines_sameslot", vec, [(l, 0) for l in range(32)])


if __name__ == "__main__":
    sys.exit(main())
