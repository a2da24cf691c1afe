"""CPU model of liveness-based row numbering for shortened tapes (the design the
models with renamed slots need next).  Today a walker of a tape with more than 32 slot ids renames ids to
shared-memory rows on first sight (tape_stream.cuh, TapeStream::rename): one row per distinct id.  A tape
push walks its tape BACKWARDS holding the set of live slots (context.cu:323-458, k_eval_tiles in
kernels.cu), which is a linear-scan register allocator run in reverse: a value takes a row at its last
reader and gives it back at the clause that defines it, so the rows in use never exceed the values live
at once.  This file states that allocator and checks it: walking a tape by rows computes what walking it
by slot ids computes.  Test infrastructure only (tests/test_host.py)."""
import numpy as np

# which operands an opcode reads (mpr::Opcode values, common.cuh; same classes as tools/gen_float_loop.py)
USES_L = {2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 24, 26, 28}
USES_R = {14, 16, 18, 20, 22, 23, 25, 26, 29}


def fields(w):
    w = int(w)
    return w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, (w >> 24) & 0xff


def split(flat):
    """A flattened logical tape (oracle.tape_flatten): header cell, clauses, end cell."""
    hdr = fields(flat[0])
    body = [c for c in flat[1:] if fields(c)[0] > 1]
    end = [c for c in flat[1:] if fields(c)[0] == 0][-1]
    return hdr[1:], body, fields(end)[1]


def allocate(flat):
    """Backward linear scan.  Returns (rows of the three axes, [(row_out, row_lhs, row_rhs)] per clause,
    row of the result, rows used, values live at once at most).  Row 0 means "no operand"."""
    axes, body, result = split(flat)
    free, next_row, row_of = [], 1, {}

    def take():
        nonlocal next_row
        if free:
            return free.pop()
        next_row += 1
        return next_row - 1

    row_of[result] = take()
    result_row = row_of[result]
    rows, most = [None] * len(body), 1
    for i in range(len(body) - 1, -1, -1):
        op, o, l, r = fields(body[i])
        most = max(most, len(row_of) + (0 if o in row_of else 1))      # live behind this clause (its result included)
        ro = row_of.pop(o, None)
        if ro is None:                      # a value nobody reads (the push never emits one): a scratch row
            ro = take()
        free.append(ro)                     # the operands whose last reader this clause is may take it over
        rl = rr = 0
        if op in USES_L and l:
            if l not in row_of:
                row_of[l] = take()
            rl = row_of[l]
        if op in USES_R and r:
            if r not in row_of:
                row_of[r] = take()
            rr = row_of[r]
        rows[i] = (ro, rl, rr)
        most = max(most, len(row_of))
    # what is still live in front of the first clause can only be an axis
    assert set(row_of) <= {a for a in axes if a}, (sorted(row_of), axes)
    axis_rows = tuple(row_of.get(a, 0) if a else 0 for a in axes)
    return axis_rows, rows, result_row, next_row - 1, most


def clause(op, l, r, imm):
    """Float semantics of one clause, enough to tell values apart (context.cu:887-920)."""
    f = np.float32
    with np.errstate(all="ignore"):
        if op == 2: return f(l * l)
        if op == 3: return f(np.sqrt(np.abs(l)))
        if op == 4: return f(-l)
        if op in (5, 6, 7, 8, 9): return f(np.sin(l) + op)
        if op == 10: return f(np.exp(np.clip(l, -20, 20)))
        if op == 11: return f(np.abs(l))
        if op == 12: return f(np.log(np.abs(l) + f(1)))
        if op == 13: return f(l + imm)
        if op == 14: return f(l + r)
        if op == 15: return f(l * imm)
        if op == 16: return f(l * r)
        if op == 17: return f(min(l, imm))
        if op == 18: return f(min(l, r))
        if op == 19: return f(max(l, imm))
        if op == 20: return f(max(l, r))
        if op == 21: return f(l - imm)
        if op == 22: return f(imm - r)
        if op == 23: return f(l - r)
        if op == 24: return f(l / imm) if imm else f(0)
        if op == 25: return f(imm / r) if r else f(0)
        if op == 26: return f(l / r) if r else f(0)
        if op == 27: return f(imm)
        if op == 28: return f(l)
        if op == 29: return f(r)
    raise ValueError(op)


def run_by_ids(flat, xyz):
    axes, body, result = split(flat)
    v = {0: np.float32(0)}
    for a, x in zip(axes, xyz):
        if a:
            v[a] = np.float32(x)
    for c in body:
        op, o, l, r = fields(c)
        imm = np.frombuffer(np.uint32(int(c) >> 32).tobytes(), dtype=np.float32)[0]
        v[o] = clause(op, v.get(l, np.float32(0)), v.get(r, np.float32(0)), imm)
    return v[result]


def run_by_rows(flat, xyz):
    axes, body, _ = split(flat)
    axis_rows, rows, result_row, n_rows, _ = allocate(flat)
    v = np.zeros(n_rows + 1, dtype=np.float32)
    for ar, x in zip(axis_rows, xyz):
        if ar:
            v[ar] = np.float32(x)
    for c, (ro, rl, rr) in zip(body, rows):
        op = fields(c)[0]
        imm = np.frombuffer(np.uint32(int(c) >> 32).tobytes(), dtype=np.float32)[0]
        v[ro] = clause(op, v[rl], v[rr], imm)
    return v[result_row]


def first_sight_rows(flat):
    """Rows the current renaming hands out: one per distinct slot id, axes first."""
    axes, body, _ = split(flat)
    seen = []
    for a in axes:
        if a and a not in seen:
            seen.append(a)
    for c in body:
        for x in fields(c)[1:]:
            if x and x not in seen:
                seen.append(x)
    return len(seen)
