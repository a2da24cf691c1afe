"""CPU model of what k_eval_root / k_eval_sub (mpr_b200/csrc/kernels.cu) do to a tape, next to the
reference's own backward slot walk (context.cu:323-458), on arbitrary min / max verdicts - no interval
arithmetic involved.  Used by tests/test_host.py: the set of clauses a child keeps must be the same
whether it is found by walking the parent's shortened tape backwards slot by slot (reference) or by
marking the parent's dependency-level plan top-down (k_eval_sub on the plan k_eval_root wrote)."""
import random

MINLI, MAXLR, COPY_IMM, COPY_LHS, COPY_RHS = 17, 20, 27, 28, 29     # mpr::Opcode values (common.cuh)


def fields(w):
    return w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, (w >> 24) & 0xff


def is_choice(op):
    return MINLI <= op <= MAXLR


def reference_push(tape, result_slot, verdict):
    """The reference's push: walk backwards with a set of live SLOTS.  Returns (shortened tape, indices
    of the kept clauses, indices of all clauses met with a live output slot)."""
    live = {result_slot}
    out, kept, seen = [], [], []
    for i in range(len(tape) - 1, -1, -1):
        w = tape[i]
        op, o, l, r = fields(w)
        if o not in live:
            continue
        c = verdict[i] if is_choice(op) else 0
        live.discard(o)
        seen.append(i)
        e, emit = w, True
        if c == 0:
            if l:
                live.add(l)
            if r:
                live.add(r)
        elif c == 1:
            live.add(l)
            if l == o:
                emit = False
            else:
                e = (w & ~0xff) | COPY_LHS
        elif r:
            live.add(r)
            if r == o:
                emit = False
            else:
                e = (w & ~0xff) | COPY_RHS
        else:
            e = (w & ~0xff) | COPY_IMM
        if emit:
            out.append(e)
            kept.append(i)
    return out[::-1], kept[::-1], seen[::-1]


def root_plan(tape, axes):
    """api.cu build_root_plan: SSA sources, previous occupant of each output slot, dependency depth.
    Value ids: ('ax', k) for an axis, an int for clause i, None for no operand."""
    writer = {a: ("ax", k) for k, a in enumerate(axes) if a}
    lsrc, rsrc, prevw, depth = [], [], [], []
    dep = lambda v: depth[v] if isinstance(v, int) else 0
    for i, w in enumerate(tape):
        op, o, l, r = fields(w)
        lsrc.append(writer.get(l) if l else None)
        rsrc.append(writer.get(r) if r else None)
        prevw.append(writer.get(o))
        depth.append(1 + max(dep(lsrc[-1]), dep(rsrc[-1])))
        writer[o] = i
    return lsrc, rsrc, prevw, depth, writer


def planned_child_live(tape, axes, result_slot, verdict0, verdict1_of_cell):
    """k_eval_root on `tape` with verdict0 (mark on SSA ids, plan emission), then k_eval_sub's mark on that
    plan with the child's verdicts (given per cell of the shortened tape).  Returns (kept root indices of the
    parent's push, set of shortened-tape cells the child finds live, number of top-down sweeps it took)."""
    n = len(tape)
    lsrc, rsrc, prevw, depth, writer = root_plan(tape, axes)
    result = writer[result_slot]
    live = {result}
    for i in range(n - 1, -1, -1):                  # any order that visits consumers first will do here
        if i not in live:
            continue
        op = tape[i] & 0xff
        c = verdict0[i] if is_choice(op) else 0
        for v in ((lsrc[i], rsrc[i]) if c == 0 else (lsrc[i],) if c == 1 else (rsrc[i],)):
            if v is not None:
                live.add(v)
    kept, in_tape = [], set()
    for i in range(n):
        if i not in live:
            continue
        op, o, l, r = fields(tape[i])
        c = verdict0[i] if is_choice(op) else 0
        if not ((c == 1 and l == o) or (c == 2 and r != 0 and r == o)):
            kept.append(i)
            in_tape.add(i)
    cell_of = {i: q for q, i in enumerate(kept)}

    def in_plan(v):                                 # root value -> the shortened tape's last writer of its slot
        while isinstance(v, int) and v not in in_tape:
            v = prevw[v]
        return v

    plan = {}                                       # kept root index -> (copy?, lsrc, rsrc) in plan values
    for i in kept:
        op, o, l, r = fields(tape[i])
        c = verdict0[i] if is_choice(op) else 0
        if c == 0:
            plan[i] = (False, in_plan(lsrc[i]), in_plan(rsrc[i]))
        elif c == 1 or r != 0:
            chosen, other = (lsrc[i], rsrc[i]) if c == 1 else (rsrc[i], lsrc[i])
            plan[i] = (True, in_plan(chosen), in_plan(other))
        else:
            plan[i] = (True, None, in_plan(lsrc[i]))            # COPY_IMM: only the stale operand is marked
    # ---- k_eval_sub: top-down by the ROOT's levels, restarting when an edge marked a level already swept ----
    by_level = sorted(kept, key=lambda i: -depth[i])
    marked = {in_plan(result)}
    sweeps, top = 0, max(depth) + 1
    while top is not None:
        sweeps += 1
        late = None
        for i in by_level:
            if depth[i] > top or i not in marked:
                continue
            copy, a, b = plan[i]
            op = tape[i] & 0xff
            c = 0 if copy or not is_choice(op) else verdict1_of_cell[cell_of[i]]
            if copy and isinstance(b, int) and b not in marked and depth[b] >= depth[i]:
                late = depth[b] if late is None else max(late, depth[b])
            for v in ((a, b) if c == 0 else (a,) if c == 1 else (b,)):
                if v is not None:
                    marked.add(v)
        top = late
    return kept, {cell_of[i] for i in marked if isinstance(i, int)}, sweeps


def compare(cells, seed, p_root, p_child):
    """cells: packed tape (uint64 array incl. header and end cell).  Random verdicts with the given
    probabilities of being decided.  Returns (reference live cells, planned live cells, sweeps)."""
    rng = random.Random(seed)
    n = len(cells) - 2
    hdr = int(cells[0]) & 0xffffffff
    axes = [(hdr >> 8) & 0xff, (hdr >> 16) & 0xff, hdr >> 24]
    tape = [int(cells[i]) & 0xffffffff for i in range(1, n + 1)]
    result_slot = (int(cells[n + 1]) >> 8) & 0xff
    v0 = [rng.choice((1, 2)) if rng.random() < p_root else 0 for _ in tape]
    short, kept_ref, _ = reference_push(tape, result_slot, v0)
    v1 = [rng.choice((1, 2)) if rng.random() < p_child else 0 for _ in short]
    _, _, seen = reference_push(short, result_slot, v1)
    kept, mine, sweeps = planned_child_live(tape, axes, result_slot, v0, v1)
    assert kept == kept_ref
    return set(seen), mine, sweeps
