"""CPU oracle for the route -> ensure-resident -> predict path of mKaloer/TFServingCache.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU baseline.  The
product (``tfservingcache_b200`` + ``libtfsc_b200.so``) never imports it and has no CPU
fallback.

Parity pin status:
  * LRU (``oracle.lrucache``)            -- PINNED by the reference's lrucache_test.go vectors.
  * URL / version parsing, 404 / 400     -- PINNED by tfservingproxy_test.go.
  * disk version-dir matching            -- PINNED by diskmodelprovider_test.go.
  * ring placement (``oracle.ring``)     -- the hash ring is the un-vendored module
    stathat.com/c/consistent v1.0.0 (go.mod:25); the reference's own tests pin properties only
    (determinism, single member, restore after membership change). PINNED ON RECALLED VECTORS: the
    restatement reproduces every expectation of the module's own test-suite (consistent_test.go:
    TestGetMultiple, TestGetMultipleRemove, TestGetTwo, TestGetN, TestGetNLess, TestGetNMore --
    tests/test_oracle_pins.py), but those vectors are quoted from memory of the public source, not
    read from a copy of the module; with the module itself absent this is the strongest pin available.
  * Predict numerics (``oracle.models``) -- PARITY UNPINNED: arithmetic lives in an external,
    unpinned tensorflow/serving image (deploy/docker-compose/docker-compose.yaml:23).  The only
    known answer in the reference is half_plus_two [1,2,5] -> [2.5,3.0,4.5]
    (deploy/docker-compose/readme.md:40-42), which the oracle reproduces.
"""
