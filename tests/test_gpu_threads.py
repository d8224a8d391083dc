"""The C ABI under threads (the Rust side may call a handle from rayon workers, src/runner.rs:388-416):
calls on ONE handle serialise on the handle's lock, different handles run concurrently on their own streams."""
import threading

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _state(g):
    off, pos, base = g.export_mutations()
    return off, pos, base, g.get_raw_fitness().view(np.uint64)


def test_threads_share_one_handle():
    from test_gpu_stochastic import big_singlehost
    c = big_singlehost(n=60000, mu=1e-4, seed=9)
    a, b = helpers.make_gpu(c), helpers.make_gpu(c)
    for g in (a, b):
        g.run(2)
        g.increment_generation()
        g.infect()
        g.mutate_infectants()
        g.replicate_infectants(to_host=False)
    ts = a.target_size(None)
    T, K = 8, 10
    sizes, errors = [[] for _ in range(T)], []

    def worker(t):
        try:
            for _ in range(K):
                p = a.subsample_population(None, 0.01)
                sizes[t].append(len(p))
                p.free()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert all(s == int(0.01 * ts) for row in sizes for s in row)
    for _ in range(T * K):  # the same number of calls, one after the other
        b.subsample_population(None, 0.01).free()
    # both handles have consumed the same draw streams: the next generation is the same
    for g in (a, b):
        g.set_population(g.subsample_population(None, 1.0))
        g.run(2)
    for x, y in zip(_state(a), _state(b)):
        assert np.array_equal(x, y)
    a.close()
    b.close()


def test_threads_drive_their_own_handles():
    from test_gpu_stochastic import big_singlehost
    T = 6
    cases = [big_singlehost(n=50000 + 1000 * t, mu=1e-4, seed=100 + t) for t in range(T)]
    serial = []
    for c in cases:
        g = helpers.make_gpu(c)
        g.run(6)
        serial.append(_state(g))
        g.close()
    out, errors = [None] * T, []

    def worker(t):
        try:
            g = helpers.make_gpu(cases[t])
            for _ in range(3):
                g.run(2)
            out[t] = _state(g)
            g.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(T):
        for x, y in zip(out[t], serial[t]):
            assert np.array_equal(x, y), f"handle {t}"
