"""Oracle restatement of TrioCaller against the reference suite (TrioCallerSuite.scala:84-237).
Alleles: 0 REF, 1 ALT, 2 OTHER_ALT, 3 NO_CALL."""
R, A, O, N = 0, 1, 2, 3


def test_filter_ref(oracle):                                            # :84-107
    assert not oracle.trio_process(None, None, None)["kept"]
    assert not oracle.trio_process([R, R], None, None)["kept"]
    assert oracle.trio_process([R, R], [R, A], None)["kept"]


def test_missing_parents_are_filled_with_no_calls(oracle):              # :109-131
    r = oracle.trio_process(None, None, [R, A])
    assert r["kept"] and r["filled"] == [True, True, False] and r["childAlleles"] == [R, A] and r["childAction"] == 0


def test_odd_copy_number_passes_through(oracle):                        # :133-153
    r = oracle.trio_process([R], [R, A], [R, A])
    assert r["kept"] and r["childAction"] == 0 and r["childAlleles"] == [R, A] and r["filled"] == [False] * 3
    assert oracle.trio_process([R, O], [R, A], [R, A])["childAction"] == 0
    assert oracle.trio_process([R, N], [R, A], [R, A])["childAction"] == 0


def test_consistent_and_phased(oracle):                                 # :155-183
    r = oracle.trio_process([R, R], [R, A], [A, R])
    assert r["childAction"] == 4 and r["childAlleles"] == [R, A]
    r = oracle.trio_process([R, A], [R, R], [R, A])
    assert r["childAction"] == 3 and r["childAlleles"] == [A, R]
    assert oracle.trio_process([A, A], [R, A], [A, A])["childAction"] == 1       # hom alt confirmed: phased as is
    assert oracle.trio_process([R, R], [R, A], [R, R])["childAction"] == 1


def test_consistent_but_unphasable(oracle):                             # :185-212
    r = oracle.trio_process([R, A], [R, A], [A, R])
    assert r["childAction"] == 0 and r["childAlleles"] == [A, R]


def test_inconsistent_call_is_invalidated(oracle):                      # :214-237
    r = oracle.trio_process([R, R], [R, A], [A, A])
    assert r["childAction"] == 2 and r["childAlleles"] == [N, N] and r["kept"]     # parent2 still carries ALT
    r = oracle.trio_process([R, R], [R, R], [R, A])
    assert r["childAction"] == 2 and not r["kept"]                                  # nothing variant is left
    assert oracle.trio_process([A, A], [R, A], [R, R])["childAction"] == 2
