"""The credit-free landing-buffer rule of the peer-memory slab exchange (cabanapic_b200/csrc/cpic_mgpu.cuh), as an
executable model (CPU).

A rank S stores a message of phase p straight into a landing buffer X of its neighbour R (k_p2p_put) and raises R's
arrival flag; R consumes X right after its wait of phase p.  There are no credits: S may write X again in a later
phase q (possibly of the next step) only if R is certain to have consumed the earlier message by then.  What makes it
certain: a phase t strictly between p and q in which R sends a message to S -- R's put(t) is enqueued behind R's
consumption of phase p on R's stream, and S's put(q) is enqueued behind S's wait(t).  (On 8 GPUs the first version let
the J and the cB ghost planes share their landing slots -- consecutive phases -- and landed planes in the field arrays
themselves; 13 of 3.4e7 migrations differed.  This model rejects that layout.)

The table below restates Mgpu::slab_step: phases in stream order, who sends to whom, and which landing buffer of the
RECEIVER each message goes to; tests/test_mgpu.py checks the real thing on GPUs, with one rank held back before every
send."""

# phase -> {direction: [landing buffers written in the receiver's mailbox]}; "up" = to the rank above (lands "from below")
PHASES = [
    ("P1 accumulator planes + leaver counts + leaver records", {"up": ["scratch_lo", "cnt_from_below", "recv_from_below"],
                                                                "down": ["scratch_hi", "cnt_from_above", "recv_from_above"]}),
    ("P2 cB ghost planes (after the first advance_b)", {"up": ["copy_cB_from_below"], "down": ["copy_cB_from_above"]}),
    ("P3 J fold planes (upward only)", {"up": ["scratch_lo"]}),      # rx, ry alias the low accumulator scratch plane
    ("P4 J ghost planes", {"up": ["copy_J_from_below"], "down": ["copy_J_from_above"]}),
    ("P5 cB ghost planes (after the second advance_b)", {"up": ["copy_cB_from_below"], "down": ["copy_cB_from_above"]}),
]


def violations(phases):
    """every (buffer, p, q) whose second write is not protected by a reverse message strictly between the two writes"""
    n = len(phases)
    bad = []
    for d, rev in (("up", "down"), ("down", "up")):
        writes = {}
        for i, (_, msgs) in enumerate(phases):
            for b in msgs.get(d, []):
                writes.setdefault(b, []).append(i)
        for b, ws in writes.items():
            for j, p in enumerate(ws):
                q = ws[(j + 1) % len(ws)]
                q = q if q > p else q + n                       # the next write, possibly in the next step
                between = [t % n for t in range(p + 1, q)]
                if not any(rev in phases[t][1] for t in between):
                    bad.append((b, phases[p][0][:2], phases[q % n][0][:2]))
    return bad


def test_every_landing_buffer_is_reused_safely():
    assert violations(PHASES) == []


def test_the_model_rejects_the_first_layout():
    """J and cB ghost planes sharing one set of landing slots: P4 -> P5 are consecutive phases"""
    shared = [(name, {d: ["copy_from_below" if "from_below" in b and b.startswith("copy") else
                          "copy_from_above" if "from_above" in b and b.startswith("copy") else b for b in bs]
                      for d, bs in msgs.items()}) for name, msgs in PHASES]
    bad = violations(shared)
    assert ("copy_from_below", "P4", "P5") in bad and ("copy_from_above", "P4", "P5") in bad


def test_the_model_needs_the_reverse_message():
    """were the J fold planes sent downward only, the upper neighbour's low scratch plane would still be safe, but a
    phase with no upward traffic between two downward writes of one buffer is rejected"""
    phases = [("P1", {"up": ["a"], "down": ["b"]}), ("P2", {"down": ["b"]})]      # b written twice, nothing comes back between
    assert ("b", "P1", "P2") in violations(phases)


def test_source_uses_separate_slots_for_cB_and_J():
    """the property the model relies on, read off the source: copy_slot() takes the plane kind"""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cabanapic_b200", "csrc", "cpic_mgpu.cuh")).read()
    assert re.search(r"copy_slot\(char\* mailbox, const MailLayout& L, int kind, int from, int i\)", src)
    assert "const int kind = m0 == F_JFX ? 1 : 0;" in src
    assert "12 * m->copy_stride()" in src                           # 2 kinds x 2 directions x 3 members
