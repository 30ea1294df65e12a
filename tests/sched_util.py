"""Test helper: issue a plan's launches in the most hostile order its scheduling hints allow (graph._Ops)."""


def adversarial_run(ops, stream, nl=2):
    """Issue `ops` in the most hostile order keras_api._run_ops' hints allow: launches without a lane are barriers;
    between two barriers every stream (lane % nl) keeps its own order and `chain` launches keep their emission order,
    but otherwise the HIGHEST stream always goes first.  Side launches are delayed to the next `join` / the end."""
    i, n = 0, len(ops)
    delayed = []
    while i < n:
        if getattr(ops[i], "lane", None) is None and not getattr(ops[i], "side", False):
            if getattr(ops[i], "join", False):
                for d in delayed:
                    d(stream)
                delayed.clear()
            ops[i](stream)
            i += 1
            continue
        j = i
        while j < n and (getattr(ops[j], "lane", None) is not None or getattr(ops[j], "side", False)):
            j += 1
        region = ops[i:j]
        queues = {}
        for k, op in enumerate(region):
            if getattr(op, "side", False):
                delayed.append(op)
            else:
                queues.setdefault(op.lane % nl, []).append((k, op))
        chain_pos = {}
        for k, op in enumerate(region):
            ck = getattr(op, "chain", None)
            if ck is not None and not getattr(op, "side", False):
                chain_pos.setdefault(ck, []).append(k)
        done = set()
        while any(queues.values()):
            progressed = False
            for s in sorted(queues, reverse=True):
                q = queues[s]
                while q:
                    k, op = q[0]
                    ck = getattr(op, "chain", None)
                    if ck is not None and any(p < k and p not in done for p in chain_pos[ck]):
                        break
                    assert not getattr(op, "join", False), "a join inside a lane region is not expected"
                    op(stream)
                    done.add(k)
                    q.pop(0)
                    progressed = True
                if progressed:
                    break               # restart from the highest stream
            assert progressed, "hints deadlock"
        i = j
    for d in delayed:
        d(stream)
