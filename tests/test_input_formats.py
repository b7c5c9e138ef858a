"""Read-input edge cases of the bowtie-compatible driver, checked against the reference binary on the same files.

The reference's parsers (pat.cpp: FASTQ 797-975, FASTA 531-640, raw 1129-1213, -c 357-523) work on characters, not on lines,
in two steps — a "light parse" that cuts the file into records and parse() that picks name/sequence/qualities out of a
record — and the cases below are the places where that shows: blank lines, CRs, '>' or '+' in the wrong place, files that end
inside a record (which also costs the reference the record BEFORE the incomplete one), records that close a light-parse batch
of 16, reads too short to search.  Output, exit status and the messages on stderr must all match.  Host logic only (the search
runs on the emulation shim), so this runs in the CPU suite.
"""
import os
import subprocess
from pathlib import Path

import pytest

from helpers import FIXTURES, REF_ALIGN, ROOT, ensure_oracle_built, have_reference
from test_cli_parity import SHIM_DIR, build_shim

CLI = ROOT / "bowtie_b200" / "bowtie-b200-align"


@pytest.fixture(scope="module")
def cli():
    ensure_oracle_built()
    if not have_reference():
        pytest.skip("reference binary / fixtures not available")
    import bowtie_b200
    bowtie_b200.build_library()
    build_shim()
    return CLI


def recs(n, start=0):
    lines = (FIXTURES / "e_coli_1000.fq").read_text().splitlines()
    return [lines[i:i + 4] for i in range(4 * start, 4 * (start + n), 4)]


def both(cli, tmp_path, tag, flags, src):
    """src: list of arguments naming the reads (files are created by the caller).  Returns nothing; asserts equality."""
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    res = []
    for who, exe, extra, e in (("ref", REF_ALIGN, ["-p", "1"], None), ("our", cli, [], env)):
        out = tmp_path / f"{tag}.{who}"
        p = subprocess.run([str(exe), *flags, *extra, "-x", str(FIXTURES / "e_coli"), *src, str(out)], capture_output=True, text=True, env=e)
        msgs = [l for l in p.stderr.splitlines() if not l.startswith("Command:")]
        body = out.read_bytes() if p.returncode == 0 and out.exists() else b""
        res.append((p.returncode, body, msgs))
    assert res[0][0] in (0, 1), (tag, res[0])            # a reference crash is not a specification
    assert res[0] == res[1], tag


def fq(rs):
    return "".join("\n".join(r) + "\n" for r in rs)


def fa(rs):
    return "".join(f">{r[0][1:]}\n{r[1]}\n" for r in rs)


def raw(rs):
    return "".join(r[1] + "\n" for r in rs)


X = recs(1, 20)[0]          # a record to doctor


def fastq_cases():
    name, seq, plus, qual = X
    one = lambda n_, s_, p_, q_: "\n".join([n_, s_, p_, q_]) + "\n"
    for n in (3, 15, 16, 17, 32):
        b = fq(recs(n))
        yield f"ok{n}", b
        yield f"nonl{n}", b.rstrip("\n")
        if n % 16:      # (in the first slot of a light-parse batch the reference goes on to parse leftovers of its buffers: not pinned here)
            yield f"trail1_{n}", b + "\n"
            yield f"trail2_{n}", b + "\n\n"
            yield f"cut1_{n}", b + "@x\n"
            yield f"cut2_{n}", b + "@x\nACGT\n"
        # (three trailing newlines make the reference parse a record of stale bytes: its message varies from run to run)
        yield f"cut0_{n}", b + "@x"
        yield f"cut3_{n}", b + "@x\nACGTACGTACGTACGTACGTAACCGT\n+\n"
    b = fq(recs(5))
    yield "crlf", b + one(name, seq, plus, qual).replace("\n", "\r\n") + b
    yield "blank", b + "\n" + b
    yield "blankname", b + one(name + "\n", seq, plus, qual) + b
    yield "seq2lines", b + one(name, seq[:10] + "\n" + seq[10:], plus, qual) + b
    yield "dot", b + one(name, seq[:5] + "." + seq[6:], plus, qual) + b
    yield "dash", b + one(name, seq[:7] + "-" + seq[7:], plus, qual) + b
    yield "plusinseq", b + one(name, seq[:7] + "+" + seq[7:], plus, qual) + b
    yield "lower", b + one(name, seq.lower(), plus, qual) + b
    yield "noname", b + one("@", seq, plus, qual) + b
    yield "crname", b + one(name[:3] + "\r" + name[3:], seq, plus, qual) + b
    yield "space", b + one(name, seq, plus, qual[:4] + " " + qual[5:]) + b
    yield "space0", b + one(name, seq, plus, " " + qual[1:]) + b
    yield "shortq", b + one(name, seq, plus, qual[:-2]) + b
    yield "longq", b + one(name, seq, plus, qual + "II") + b
    yield "noat", b + one("X" + name[1:], seq, plus, qual) + b
    yield "lowq", b + one(name, seq, plus, qual[:3] + "\x1f" + qual[4:]) + b
    yield "plusname", b + one(name, seq, "+" + name[1:], qual) + b
    yield "shortread", b + one(name, "ACG", plus, "III") + b
    yield "emptyread", b + one(name, "", plus, "") + b
    yield "empty", ""
    yield "newline", "\n"
    yield "notfastq", "ACGT\n"
    yield "leading", "\n\r\n" + b


def test_fastq_edge_cases(cli, tmp_path):
    for tag, text in fastq_cases():
        f = tmp_path / f"fq_{tag}.fq"
        f.write_text(text)
        both(cli, tmp_path, f"fq_{tag}", ["-n", "2"], ["-q", str(f)])
    # trimming and quality encodings go through the same record parser
    f = tmp_path / "fq_ok17.fq"
    for i, flags in enumerate((["-5", "3", "-3", "2"], ["--phred64-quals"], ["--solexa-quals"], ["-5", "40"], ["-3", "33"])):
        both(cli, tmp_path, f"fq_flags{i}", ["-n", "2", *flags], ["-q", str(f)])


def fasta_cases():
    name, seq = X[0][1:], X[1]
    for n in (3, 16, 17):
        b = fa(recs(n))
        yield f"ok{n}", b
        yield f"nonl{n}", b.rstrip("\n")                       # the last base is lost (pat.cpp:607-619)
        yield f"trail{n}", b + "\n\n"
        yield f"gt{n}", b + ">"
        yield f"gtname{n}", b + ">x\n"
        yield f"gtname_nonl{n}", b + ">x"
        yield f"onebase{n}", b + ">x\nA"
        yield f"multi{n}", f">m\n{seq[:20]}\n{seq[20:]}\n" + b   # only the first line is the read
        yield f"blank{n}", f">m\n\n\n{seq}\n" + b
        yield f"gtinname{n}", f">m>k\n{seq}\n" + b               # '>' ends a record wherever it is
        yield f"noname{n}", f">\n{seq}\n" + b
        yield f"crlf{n}", b.replace("\n", "\r\n")
        yield f"dotdash{n}", f">m\n{seq[:5]}.-{seq[5:]}\n" + b
        yield f"lead{n}", "\n\r\n" + b
        yield f"emptyseq{n}", ">m\n>k\n" + b
    yield "empty", ""
    yield "newline", "\n"
    yield "notfasta", "ACGT\n"


def test_fasta_edge_cases(cli, tmp_path):
    for tag, text in fasta_cases():
        f = tmp_path / f"fa_{tag}.fa"
        f.write_text(text)
        both(cli, tmp_path, f"fa_{tag}", ["-n", "2"], ["-f", str(f)])
    f = tmp_path / "fa_multi17.fa"
    both(cli, tmp_path, "fa_dump", ["-n", "2", "--un", str(tmp_path / "fa_un.fa"), "--al", str(tmp_path / "fa_al.fa")], ["-f", str(f)])


def raw_cases():
    seq, seq2 = X[1], recs(1, 21)[0][1]
    for n in (3, 16, 17):
        b = raw(recs(n))
        yield f"ok{n}", b
        yield f"nonl{n}", b.rstrip("\n")
        yield f"blank{n}", "\n\n" + b + "\n\n"
        yield f"digits{n}", "12345\n" + b                      # an empty read, not a skipped line
        yield f"dotdash{n}", f"{seq[:5]}.-{seq[5:]}\n" + b       # '.' is not N in this format
        yield f"cr{n}", f"{seq}\r{seq2}\n" + b                   # a CR separates records too
        yield f"crlf{n}", b.replace("\n", "\r\n")
    yield "empty", ""


def test_raw_edge_cases(cli, tmp_path):
    for tag, text in raw_cases():
        f = tmp_path / f"raw_{tag}.txt"
        f.write_text(text)
        both(cli, tmp_path, f"raw_{tag}", ["-n", "2"], ["-r", str(f)])


def test_cmdline_reads_edge_cases(cli, tmp_path):
    s0, q0 = X[1], X[3]
    s1 = recs(1, 6)[0][1]
    cases = [([s0], []), ([s0 + ":" + q0], []), ([s0 + ":" + q0[:-1]], []), ([s0 + ":" + q0 + "I"], []), ([s0[:5] + "." + s0[5:]], []), ([s0 + ":"], []),
             ([s0, s1], []), ([s0 + ":" + q0.replace(q0[3], " ")], []), ([s0 + ":" + q0], ["--solexa-quals"]), ([s0], ["-5", "3", "-3", "2"]),
             ([s0 + ":" + q0], ["-5", "3", "-3", "2"])]
    for i, (toks, flags) in enumerate(cases):
        both(cli, tmp_path, f"c_{i}", ["-n", "2", *flags], ["-c", ",".join(toks)])


SHORT_MODES = [["-v", "0"], ["-v", "1"], ["-v", "2"], ["-v", "3"], ["-n", "0"], ["-n", "2"], ["-n", "3"], ["-n", "2", "--best"], ["-v", "1", "--best"],
               ["-n", "2", "--quiet"], ["-v", "3", "--quiet"]]


def test_reads_too_short_to_search(cli, tmp_path):
    """search_1mm_phase1.c:12-15, search_23mm_phase1.c:13-20 (errors), search_seeded_phase1.c:17-21, aligner.h:440-448 (warnings); --quiet
    silences the warnings but not the summary (hit.h:160)."""
    good = X[1]
    k = 0
    for mode in SHORT_MODES:
        for short in ("A", "AC", "ACG", "ACGT"):
            for order in (short + "," + good, good + "," + short):
                k += 1
                both(cli, tmp_path, f"s_{k}", mode, ["-c", order])
    m1, m2 = tmp_path / "sp_1.fq", tmp_path / "sp_2.fq"
    r = recs(4)
    m1.write_text(fq(r[:2]) + "@p\nACG\n+\nIII\n" + fq(r[2:]))
    m2.write_text(fq(r[2:]) + "@p\nACGTACGTAGGCTAGCTAGGATCGATTTAGGCAT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n" + fq(r[:2]))
    both(cli, tmp_path, "s_pair", ["-n", "2"], ["-1", str(m1), "-2", str(m2)])


def test_fasta_continuous(cli, tmp_path):
    """-F <len>,<freq> (FastaContinuousPatternSource, pat.cpp:651-790): reads are windows of the FASTA input, named after the
    sequence and the offset; IUPAC codes and '-' are N, other characters are skipped, trimming options do not apply."""
    genome = None
    for cand in (ROOT / "oracle" / "_ref" / "fixtures" / "NC_008253.fna", Path("/root/reference/genomes/NC_008253.fna")):
        if cand.exists():
            genome = cand.read_text()
            break
    if genome is None:
        pytest.skip("no genome FASTA to cut windows from")
    g1, g2 = tmp_path / "g1.fa", tmp_path / "g2.fa"
    g1.write_text(genome[:6000])
    body = genome[100:1500].split("\n", 1)[1][:700]
    g2.write_text(">seqA some words\n" + body + "\n>seqB\nACGTNNRYACGTTTGACCAGT-ACGATGCAGCTAGCTAGCTAGGATCGATCGATGCTAGCTAGCTAGCATGCATGCTAGTCGAT\nGGATTCAG*GATTAGCAT12ACAGGATACCAGGGATATTACAC\n")
    for lf in ("30,10", "35,1", "50,7", "100,40"):
        both(cli, tmp_path, f"F1_{lf}", ["-F", lf, "-n", "2"], [str(g1)])
        both(cli, tmp_path, f"F2_{lf}", ["-F", lf, "-v", "2"], [str(g2)])
        both(cli, tmp_path, f"F3_{lf}", ["-F", lf, "-v", "2", "-5", "2", "-3", "3"], [f"{g2},{g1}"])


def test_quality_options(cli, tmp_path):
    """--integer-quals (pat.cpp:905-923, qual.h:135-153); -Q/--Q1/--Q2 quality files are opened and counted but never read
    (pat.cpp:333-348, pat.h:334-340); --12 ignores the quality-encoding flags (ebwt_search.cpp:2942-2943)."""
    rs = recs(40)
    iq = tmp_path / "iq.fq"
    iq.write_text("".join(f"{r[0]}\n{r[1]}\n+\n{' '.join(str(ord(ch) - 33) for ch in r[3])}\n" for r in rs))
    both(cli, tmp_path, "iq", ["-n", "2", "--integer-quals"], [str(iq)])
    both(cli, tmp_path, "iq_sol", ["-n", "2", "--integer-quals", "--solexa-quals"], [str(iq)])
    both(cli, tmp_path, "iq_t5", ["-n", "2", "--integer-quals", "-5", "3"], [str(iq)])
    big = tmp_path / "iqbig.fq"
    big.write_text(iq.read_text().replace(" 30", " 130"))
    both(cli, tmp_path, "iq_big", ["-n", "2", "--integer-quals"], [str(big)])
    both(cli, tmp_path, "iq_noflag", ["-n", "2"], [str(iq)])
    rfa, rq = tmp_path / "r.fa", tmp_path / "r.qual"
    rfa.write_text(fa(rs)); rq.write_text("x\n")
    both(cli, tmp_path, "Q_ok", ["-n", "2", "-f", "-Q", str(rq)], [str(rfa)])
    both(cli, tmp_path, "Q_count", ["-n", "2", "-f", "-Q", f"{rq},{rq}"], [str(rfa)])
    both(cli, tmp_path, "Q_missing", ["-n", "2", "-f", "-Q", str(tmp_path / "none.qual")], [str(rfa)])
    both(cli, tmp_path, "Q_missing2", ["-n", "2", "-f", "-Q", f"{tmp_path / 'none.qual'},{rq}"], [f"{rfa},{rfa}"])
    tab = tmp_path / "t.tab"
    tab.write_text("".join(f"{r[0][1:]}\t{r[1]}\t{r[3]}\n" for r in rs))
    both(cli, tmp_path, "tab_p64", ["-n", "2", "--phred64-quals"], ["--12", str(tab)])
    both(cli, tmp_path, "tab_sol", ["-n", "2", "--solexa-quals"], ["--12", str(tab)])


def test_option_aliases_and_hidden_switches(cli, tmp_path):
    fq1 = str(FIXTURES / "e_coli_1000.fq")
    pair = ["-1", str(FIXTURES / "e_coli_1000_1.fq"), "-2", str(FIXTURES / "e_coli_1000_2.fq")]
    both(cli, tmp_path, "khits", ["-n", "2", "--khits", "3", "--mhits", "5"], [fq1])
    both(cli, tmp_path, "stateful", ["-n", "2", "--stateful"], [fq1])                 # best-first aligners, paired stays V1
    both(cli, tmp_path, "stateful_p", ["-n", "2", "--stateful", "-u", "200"], pair)
    both(cli, tmp_path, "pev2", ["-n", "2", "--pev2", "-u", "200"], pair)              # PairedBWAlignerV2 without --best
    both(cli, tmp_path, "offrate", ["-n", "2", "-o", "7"], [fq1])
    both(cli, tmp_path, "noop", ["-n", "2", "--strandfix", "--noreconcile", "--chunksz", "32"], [fq1])


def test_paired_inputs_edge_cases(cli, tmp_path):
    """-1/-2: DualPatternComposer::nextBatch compares what the two files delivered (pat.cpp:164-222) — a file that is short, or
    that loses its last record to a trailing blank line, is an error; --interleaved counts pairs, so a last record without a
    mate is dropped and blank lines after the last pair cost that pair; --12 records with 4 fields are skipped."""
    def L(p):
        l = (FIXTURES / p).read_text().splitlines()
        return [l[i:i + 4] for i in range(0, len(l), 4)]
    M1, M2 = L("e_coli_1000_1.fq"), L("e_coli_1000_2.fq")

    def w(name, text):
        f = tmp_path / name
        f.write_text(text)
        return str(f)
    for n in (5, 16, 17):
        a, b = fq(M1[:n]), fq(M2[:n])
        both(cli, tmp_path, f"p_ok{n}", ["-n", "2"], ["-1", w("a.fq", a), "-2", w("b.fq", b)])
        both(cli, tmp_path, f"p_short2_{n}", ["-n", "2"], ["-1", w("a.fq", a), "-2", w("b.fq", fq(M2[:n - 1]))])
        both(cli, tmp_path, f"p_short1_{n}", ["-n", "2"], ["-1", w("a.fq", fq(M1[:n - 1])), "-2", w("b.fq", b)])
        both(cli, tmp_path, f"p_trail1_{n}", ["-n", "2"], ["-1", w("a.fq", a + "\n"), "-2", w("b.fq", b)])
        both(cli, tmp_path, f"p_trail2_{n}", ["-n", "2"], ["-1", w("a.fq", a), "-2", w("b.fq", b + "\n")])
        if n % 16:
            both(cli, tmp_path, f"p_trailboth_{n}", ["-n", "2"], ["-1", w("a.fq", a + "\n"), "-2", w("b.fq", b + "\n")])
        both(cli, tmp_path, f"p_nonl_{n}", ["-n", "2"], ["-1", w("a.fq", a.rstrip("\n")), "-2", w("b.fq", b.rstrip("\n"))])
        il = "".join("\n".join(x) + "\n" + "\n".join(y) + "\n" for x, y in zip(M1[:n], M2[:n]))
        both(cli, tmp_path, f"i_ok{n}", ["-n", "2"], ["--interleaved", w("i.fq", il)])
        both(cli, tmp_path, f"i_odd{n}", ["-n", "2"], ["--interleaved", w("i.fq", il + "\n".join(M1[n]) + "\n")])
        if n % 16:
            both(cli, tmp_path, f"i_trail{n}", ["-n", "2"], ["--interleaved", w("i.fq", il + "\n")])
        both(cli, tmp_path, f"i_nonl{n}", ["-n", "2"], ["--interleaved", w("i.fq", il.rstrip("\n"))])
        tab = "".join(f"{x[0][1:]}\t{x[1]}\t{x[3]}\t{y[1]}\t{y[3]}\n" for x, y in zip(M1[:n], M2[:n]))
        both(cli, tmp_path, f"t_ok{n}", ["-n", "2"], ["--12", w("t.tab", tab)])
        both(cli, tmp_path, f"t_blank{n}", ["-n", "2"], ["--12", w("t.tab", "\n" + tab + "\n\n")])
        both(cli, tmp_path, f"t_nonl{n}", ["-n", "2"], ["--12", w("t.tab", tab.rstrip("\n"))])
        both(cli, tmp_path, f"t_crlf{n}", ["-n", "2"], ["--12", w("t.tab", tab.replace("\n", "\r\n"))])
        both(cli, tmp_path, f"t_2f{n}", ["-n", "2"], ["--12", w("t.tab", tab + "nm\tACGT\n")])
        both(cli, tmp_path, f"t_4f{n}", ["-n", "2"], ["--12", w("t.tab", tab + f"nm\t{M1[0][1]}\t{M1[0][3]}\t{M2[0][1]}\n")])
        both(cli, tmp_path, f"t_badq{n}", ["-n", "2"], ["--12", w("t.tab", tab + f"nm\t{M1[0][1]}\t{M1[0][3][:-1]}\n")])
        both(cli, tmp_path, f"pf_ok{n}", ["-n", "2", "-f"], ["-1", w("a.fa", fa(M1[:n])), "-2", w("b.fa", fa(M2[:n]))])
        both(cli, tmp_path, f"pf_short{n}", ["-n", "2", "-f"], ["-1", w("a.fa", fa(M1[:n])), "-2", w("b.fa", fa(M2[:n - 1]))])
