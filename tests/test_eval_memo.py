"""exact memoisation of the caller's repeated no-grad forwards
(itexperiments.py:464-473 evaluates twice per epoch with identical parameters).
Host logic only -- CPU tensors and a counting stand-in for the kernel call."""
import pytest
import torch

import rgb_experiment_b200.memo as M


class _Owner:          # stands in for a graph.Graph (weak-referenceable)
    pass


@pytest.fixture(autouse=True)
def _fresh():
    M.clear()
    M.set_budget_mb(64)
    for k in M.stats:
        M.stats[k] = 0
    yield
    M.clear()
    M.set_budget_mb(4096)


def _op(calls):
    def fn(x):
        calls.append(1)
        return x * 2.0 + 1.0
    return fn


BIG = 1e9   # "work" large enough to pass MIN_WORK


def test_second_identical_no_grad_call_is_served_from_the_store():
    g, calls = _Owner(), []
    f = _op(calls)
    x = torch.randn(100, 7)
    with torch.no_grad():
        a = M.cached(g, ("op",), (x,), BIG, lambda: f(x))
        b = M.cached(g, ("op",), (x.clone(),), BIG, lambda: f(x))     # a different tensor object, same values
    assert len(calls) == 1 and torch.equal(a, b) and a.data_ptr() != b.data_ptr()
    assert M.stats["hits"] == 1 and M.stats["misses"] == 1


def test_changed_values_static_args_or_owner_miss():
    g, h, calls = _Owner(), _Owner(), []
    f = _op(calls)
    x = torch.randn(50, 4)
    y = x.clone()
    y[17, 2] += 1e-3                                   # one element differs
    with torch.no_grad():
        M.cached(g, ("op", 10), (x,), BIG, lambda: f(x))
        out_y = M.cached(g, ("op", 10), (y,), BIG, lambda: f(y))
        M.cached(g, ("op", 11), (x,), BIG, lambda: f(x))
        M.cached(h, ("op", 10), (x,), BIG, lambda: f(x))
    assert len(calls) == 4 and torch.equal(out_y, y * 2.0 + 1.0)


def test_fingerprint_collision_is_caught_by_the_exact_compare():
    """Two inputs with the same sum and the same strided samples but different values elsewhere."""
    g, calls = _Owner(), []
    f = _op(calls)
    x = torch.zeros(64, 4)
    y = x.clone()
    y[0, 1], y[0, 2] = 1.0, -1.0                       # sum unchanged; flat positions 1, 2 are not sampled (step 32)
    assert M.fingerprint((x,)) == M.fingerprint((y,))
    with torch.no_grad():
        M.cached(g, ("op",), (x,), BIG, lambda: f(x))
        out = M.cached(g, ("op",), (y,), BIG, lambda: f(y))
    assert len(calls) == 2 and torch.equal(out, y * 2.0 + 1.0)


def test_grad_mode_small_work_nan_and_disabled_are_never_memoised():
    g, calls = _Owner(), []
    f = _op(calls)
    x = torch.randn(10, 3)
    M.cached(g, ("op",), (x,), BIG, lambda: f(x))                       # autograd enabled
    M.cached(g, ("op",), (x,), BIG, lambda: f(x))
    with torch.no_grad():
        M.cached(g, ("op",), (x,), 10.0, lambda: f(x))                  # below MIN_WORK
        M.cached(g, ("op",), (x,), 10.0, lambda: f(x))
        xn = x.clone()
        xn[0, 0] = float("nan")
        M.cached(g, ("op",), (xn,), BIG, lambda: f(xn))
        M.cached(g, ("op",), (xn,), BIG, lambda: f(xn))
        M.set_budget_mb(0)
        M.cached(g, ("op",), (x,), BIG, lambda: f(x))
        M.cached(g, ("op",), (x,), BIG, lambda: f(x))
    assert len(calls) == 8 and M.stats["hits"] == 0 and M.held_bytes() == 0


def test_result_edited_in_place_by_the_caller_is_dropped():
    g, calls = _Owner(), []
    f = _op(calls)
    x = torch.randn(20, 3)
    with torch.no_grad():
        a = M.cached(g, ("op",), (x,), BIG, lambda: f(x))
        a += 5.0                                        # e.g. `out += x_r` in a caller
        b = M.cached(g, ("op",), (x,), BIG, lambda: f(x))
    assert len(calls) == 2 and torch.equal(b, x * 2.0 + 1.0)


def test_store_is_bounded_in_bytes_and_dies_with_the_owner():
    M.set_budget_mb(1)                                  # 1 MiB
    g, calls = _Owner(), []
    f = _op(calls)
    xs = [torch.randn(30000) + i for i in range(8)]     # 120 kB in + 120 kB out per entry
    with torch.no_grad():
        for x in xs:
            M.cached(g, ("op",), (x,), BIG, lambda: f(x))
        assert M.held_bytes() <= 1 << 20 and M.stats["evicted"] >= 4
        M.cached(g, ("op",), (xs[-1],), BIG, lambda: f(xs[-1]))         # most recent one is still there
        assert M.stats["hits"] == 1
        big = torch.randn(200000)                       # 2 x 800 kB > budget: not even fingerprinted
        M.cached(g, ("op",), (big,), BIG, lambda: f(big))
        M.cached(g, ("op",), (big,), BIG, lambda: f(big))
        assert M.stats["hits"] == 1
        # a new owner that happens to reuse the id() of a dead one must not see its entries
        x = xs[-1]
        key_owner = _Owner()
        M.cached(key_owner, ("op",), (x,), BIG, lambda: f(x))
        ident = id(key_owner)
        del key_owner
        for _ in range(100):
            o = _Owner()
            if id(o) == ident:
                n = len(calls)
                M.cached(o, ("op",), (x,), BIG, lambda: f(x))
                assert len(calls) == n + 1
                break


def test_a_training_forward_makes_the_previous_generation_droppable():
    """Epoch loop: train forward (grad on), eval, eval.  Each epoch's eval inputs differ (the optimiser stepped);
    the store must hold two generations, not one per epoch; inputs that keep hitting never age."""
    g, calls = _Owner(), []
    f = _op(calls)
    raw = torch.randn(40, 5)                            # an aggregation of the raw features: same in every epoch
    for epoch in range(5):
        x = torch.randn(40, 5) + epoch                  # depends on the parameters: new values every epoch
        M.cached(g, ("op",), (x,), BIG, lambda: f(x))   # training forward (grad enabled): a new generation ...
        M.cached(g, ("op2",), (x,), BIG, lambda: f(x))  # ... once per training phase, however many layers pass
        with torch.no_grad():
            for _ in range(2):
                M.cached(g, ("op",), (x,), BIG, lambda: f(x))
                M.cached(g, ("op",), (raw,), BIG, lambda: f(raw))
        assert len(M._store) == min(epoch, 1) + 2       # this and the previous epoch's x + raw, nothing older
    # per epoch: x misses once and hits once; raw misses only in the first epoch
    assert M.stats["misses"] == 5 + 1 and M.stats["hits"] == 5 + 9 and M.stats["stale_dropped"] == 3


def test_inference_mode_is_not_memoised():
    g, calls = _Owner(), []
    f = _op(calls)
    with torch.inference_mode():
        x = torch.randn(10, 3)
        M.cached(g, ("op",), (x,), BIG, lambda: f(x))
        M.cached(g, ("op",), (x,), BIG, lambda: f(x))
    assert len(calls) == 2 and M.stats["hits"] == 0 and M.stats["misses"] == 0
