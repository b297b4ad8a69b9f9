"""Many-pair front end: the FASTA reader against a line-by-line restatement of
Fasta::load (reference src/fa.cpp:37-83) and the reference's input rules (src/ractip.cpp:1571-1590); on a GPU,
the whole pipeline through `python -m ractip_b200`."""
import string

import numpy as np
import pytest

STRUCT = "()[].?xle "


def fasta_load_restated(text: str):
    """Fasta::load, restated: what the reference's reader makes of `text`."""
    recs, name, seq, st = [], "", "", ""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()           # getline does not produce a line after the final newline
    for line in lines:
        first = line[0] if line else "\0"
        if first == ">":
            if name:
                recs.append((name, seq, st))
            name, seq, st = line[1:], "", ""
            continue
        if first != "\0" and first not in STRUCT:
            k = 0
            while k < len(line) and line[k] in string.ascii_letters:
                k += 1
            seq += line[:k]
        else:
            k = 0
            while k < len(line) and line[k] in STRUCT:
                k += 1
            st += line[:k]
    if name:
        recs.append((name, seq, st))
    return recs


CASES = [
    ">a\nACGU\n>b\nGGCC\n",
    ">a desc ription\nACGU\nacgu\n\n>b\nGG CC\nUU12AA\n",               # lower case, blank line, non-alpha tails
    ">a\nACGUACGU\n((....))\n>b\nGGGAAACCC\n(((...)))\n",                 # structure lines
    ">a\nACGU\n.x.|\n>b\nGG\n",                                           # structure run stops at a foreign character
    "junk before\nACGU\n>a\nAC\nGU",                                      # lines before the first header, no final newline
    ">\nACGU\n>b\nGG\n",                                                  # empty name: the record is dropped
    ">a\r\nACGU\r\n>b\r\nGG\r\n",                                         # CR LF: the CR ends the alphabetic run and stays in the name
    ">a\nxACGU\nlACGU\neACGU\nACGU\n",                                    # lines that START with x, l, e are structure lines
    "",
    ">only\n",
]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_fasta_reader_follows_the_reference(k):
    from ractip_b200 import parse_fasta
    got = [(r.name, r.seq, r.str) for r in parse_fasta(CASES[k])]
    assert got == fasta_load_restated(CASES[k])


def test_fasta_reader_random_lines():
    from ractip_b200 import parse_fasta
    rng = np.random.default_rng(7)
    alphabet = list("ACGUacgu>()[].?xle N1-\r")
    for _ in range(200):
        text = "\n".join("".join(rng.choice(alphabet, size=rng.integers(0, 12))) for _ in range(rng.integers(0, 12)))
        if rng.integers(2):
            text += "\n"
        got = [(r.name, r.seq, r.str) for r in parse_fasta(text)]
        assert got == fasta_load_restated(text), repr(text)


def test_input_rules(tmp_path):
    from ractip_b200 import input_pairs, load_fasta
    a = tmp_path / "a.fa"
    b = tmp_path / "b.fa"
    a.write_text(">a1\nACGU\n>a2\nGGCC\n>a3\nUUUU\n")
    b.write_text(">b1\nAAAA\n>b2\nCCCC\n")
    assert [r.name for r in load_fasta(a)] == ["a1", "a2", "a3"]
    p = input_pairs(a, b)
    assert [(x.name, y.name) for x, y in p] == [("a1", "b1")]                      # first record of each file
    assert [(x.name, y.name) for x, y in input_pairs(a)] == [("a1", "a2")]         # first two of one file
    assert len(input_pairs(a, b, all_pairs=True)) == 6
    assert [(x.name, y.name) for x, y in input_pairs(a, all_pairs=True)] == [("a1", "a2"), ("a1", "a3"), ("a2", "a3")]
    (tmp_path / "one.fa").write_text(">x\nACGU\n")
    with pytest.raises(ValueError, match="Format error"):
        input_pairs(tmp_path / "one.fa")
    (tmp_path / "empty.fa").write_text("ACGU\n")
    with pytest.raises(ValueError, match="Format error"):
        input_pairs(a, tmp_path / "empty.fa")
    from ractip_b200 import RpError
    with pytest.raises(RpError):
        load_fasta(tmp_path / "missing.fa")


def test_result_text_is_the_reference_layout():
    from ractip_b200 import FastaRecord, JointPrediction, PairResult, format_result
    r = PairResult(FastaRecord("s1", "ACGU"), FastaRecord("s2", "GGCC"), JointPrediction("(..)", "[[..", 1.0, -1.5, 0.25, -3.0),
                   e1s=-2.0, e2s=0.0, zscore=(-1.234567, 0.5))
    assert format_result(r).splitlines()[:6] == [">s1", "ACGU", "(..)", ">s2", "GGCC", "[[.."]
    assert format_result(r, show_energy=True).splitlines()[6] == "(E: JS= -4.25 = -1.5+0.25-3, S1+S2= -2 = -2+0)"
    assert format_result(r).splitlines()[-1] == "z-score: -1.23457, 0.5"


@pytest.mark.gpu
def test_command_line_many_pairs(tmp_path, bundled, capsys):
    """`python -m ractip_b200 --all-pairs`: one GPU batch for all pairs; each pair's text equals its own run, and the
    joint structures equal those of the pipeline fed by the oracle's matrices (tests/test_gpu_parity.py pins the
    matrices, tests/test_ip.py the programme)."""
    from ractip_b200.__main__ import main
    seqs = bundled["sequences"]
    a = tmp_path / "a.fa"
    b = tmp_path / "b.fa"
    a.write_text(f">DIS\n{seqs['DIS']}\n>Tar\n{seqs['Tar']}\n")
    b.write_text(f">DIS2\n{seqs['DIS']}\n>Tarstar\n{seqs['Tarstar']}\n")
    assert main(["--all-pairs", "-e", str(a), str(b)]) == 0
    allp = capsys.readouterr().out.strip().split("\n")
    assert len(allp) == 4 * 7
    assert main(["-e", str(a), str(b)]) == 0                      # the reference's rule: first records only
    first = capsys.readouterr().out.strip().split("\n")
    assert first == allp[:7]
    assert first[0] == ">DIS" and first[3] == ">DIS2" and first[6].startswith("(E: JS= ")
    assert set(first[2]) <= set("().[]") and len(first[2]) == len(seqs["DIS"])
    assert main([str(tmp_path / "a.fa")]) == 0                    # one file: its first two records
    one = capsys.readouterr().out.strip().split("\n")
    assert one[0] == ">DIS" and one[3] == ">Tar"


@pytest.mark.gpu
def test_front_end_matches_pair_by_pair_solve(stage, bundled, model):
    from ractip_b200 import FastaRecord, default_ip_opts, default_opts, predict, solve_joint
    seqs = bundled["sequences"]
    pairs = [(FastaRecord(a, seqs[a]), FastaRecord(b, seqs[b])) for a, b in bundled["pairs"][:4]]
    res = predict(stage, pairs, show_energy=True)
    for (a, b), r in zip(pairs, res):
        p = stage.solve_probabilities(a.seq, b.seq, default_opts())
        j = solve_joint(model, a.seq, b.seq, p, default_ip_opts(), energies=True)
        assert (r.joint.r1, r.joint.r2) == (j.r1, j.r2)
        assert abs(r.joint.e3 - j.e3) < 1e-6


@pytest.mark.gpu
def test_command_line_zscore(tmp_path, bundled, capsys):
    """--zscore: the shuffles of src/ractip.cpp:1636-1657 as ONE more GPU batch; the printed statistic is reproducible
    and follows zscore_statistic over the per-shuffle energies."""
    from ractip_b200 import (ProbabilityStage, default_ip_opts, default_opts, solve_joint, solve_ss, zscore_shuffles,
                             zscore_statistic)
    from ractip_b200.__main__ import main
    seqs = bundled["sequences"]
    f = tmp_path / "p.fa"
    f.write_text(f">Tar\n{seqs['Tar']}\n>Tarstar\n{seqs['Tarstar']}\n")
    args = ["--zscore", "12", "--num-shuffling", "6", "--seed", "3", str(f)]
    assert main(args) == 0
    out1 = capsys.readouterr().out.strip().split("\n")
    assert main(args) == 0
    assert capsys.readouterr().out.strip().split("\n") == out1
    assert out1[-1].startswith("z-score: ") and len(out1) == 7
    # the same numbers by hand
    st = ProbabilityStage()
    try:
        s1, s2 = seqs["Tar"], seqs["Tarstar"]
        p = st.solve_probabilities(s1, s2, default_opts())
        j = solve_joint(st.model, s1, s2, p, default_ip_opts(), energies=True)
        e1s = solve_ss(st.model, s1, p.bp1, default_ip_opts(), energy=True)[2]
        e2s = solve_ss(st.model, s2, p.bp2, default_ip_opts(), energy=True)[2]
        r1, r2 = zscore_shuffles(s1, s2, 6, 3, 12)
        rows = []
        for a, b, q in zip(r1, r2, st.run_dense(list(zip(r1, r2)), default_opts())):
            jj = solve_joint(st.model, a, b, q, default_ip_opts(), energies=True)
            rows.append((jj.e1 + jj.e2 + jj.e3,
                         solve_ss(st.model, a, q.bp1, default_ip_opts(), energy=True)[2] +
                         solve_ss(st.model, b, q.bp2, default_ip_opts(), energy=True)[2]))
        z = zscore_statistic(j.e1 + j.e2 + j.e3, e1s + e2s, rows)
    finally:
        st.close()
    assert out1[-1] == "z-score: " + format(z[0], ".6g") + ", " + format(z[1], ".6g")
