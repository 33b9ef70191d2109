"""Tag-set loading: the host half of ``import_tcr_info`` (reference decombine.py:593-746).

Resolves the chain, applies the tag-set / species fix-ups, reads the ``.tags`` and ``.fasta`` files and
hands them to ``dcb_tagset_build`` (which replaces the six acora automata).  File lookup follows
``read_tcr_file`` (decombine.py:187-225): working directory first, then ``tagfastadir``; the reference's
third step (download from GitHub) is replaced by the bundled copy of the same data files, because GPU
nodes have no network.
"""
import gzip
import json
import os
import sys

from . import _lib

CHAINNAMS = {"a": "alpha", "b": "beta", "g": "gamma", "d": "delta"}

_CHAIN_SPELLINGS = {
    "a": ("A", "ALPHA", "TRA", "TCRA"),
    "b": ("B", "BETA", "TRB", "TCRB"),
    "g": ("G", "GAMMA", "TRG", "TCRG"),
    "d": ("D", "DELTA", "TRD", "TCRD"),
}

_NOCHAIN = ("TCR chain not recognised. \n "
            "      Please either include (one) chain name in the file name (i.e. alpha/beta/gamma/delta),\n "
            "      or use the '-c' flag with an explicit chain option (a/b/g/d, case-insensitive).")

_BUNDLE_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tagsets.json.gz")
_bundle = None


def _bundled(name, filetype):
    global _bundle
    if _bundle is None:
        with gzip.open(_BUNDLE_PATH, "rt") as fh:
            _bundle = json.load(fh)
    try:
        return _bundle[name][filetype]
    except KeyError:
        return None


def read_tcr_text(species, tagset, chain, gene, filetype, expected_dir_name):
    """Contents of ``<species>_<tagset>_TR<CHAIN><GENE>.<filetype>`` (decombine.py:187-225)."""
    stem = "%s_%s_TR%s%s" % (species, tagset, chain.upper(), gene.upper())
    expected_file = stem + "." + filetype
    candidates = [expected_file]
    if expected_dir_name:
        candidates.append(expected_dir_name + os.sep + expected_file)
    for cand in candidates:
        if os.path.isfile(cand):
            with open(cand, "rt") as fh:
                return fh.read()
    text = _bundled(stem, filetype)
    if text is None:
        print("Cannot find following file locally or in the bundled tag sets:", expected_file)
        print("Please point Decombinator to local copies of the tag and FASTA files with the '-tfdir' flag.")
        sys.exit()
    return text


def parse_fasta_regions(text):
    """Upper-cased sequence of every FASTA record, in file order (decombine.py:690-696)."""
    regions, parts = [], None
    for raw in text.splitlines():
        if raw[:1] == ">":
            if parts is not None:
                regions.append("".join(parts).upper())
            parts = []
        elif parts is not None:
            parts.append(raw.strip())
    if parts is not None:
        regions.append("".join(parts).upper())
    return regions


def parse_tag_lines(text, half_split):
    """``get_v_tags`` / ``get_j_tags`` (decombine.py:820-866): tag, jump and the two half-tag lists."""
    seqs, jumps = [], []
    for raw in text.splitlines():
        cols = raw.split()
        if not cols:
            continue
        seqs.append(cols[0])
        jumps.append(int(cols[1]))
    half1 = [s[:half_split] for s in seqs]
    half2 = [s[half_split:] for s in seqs]
    return seqs, half1, half2, jumps


def resolve_chain(inputargs):
    """Chain letter + whether it was detected from the file name (decombine.py:601-631)."""
    in_name = [x for x in CHAINNAMS.values() if x in inputargs["infile"].lower()]
    chain_detected = 1 if len(in_name) == 1 else 0
    given = inputargs.get("chain")
    if given:
        for letter, spellings in _CHAIN_SPELLINGS.items():
            if given.upper() in spellings:
                return letter, chain_detected
        print(_NOCHAIN)
        sys.exit()
    if chain_detected:
        return in_name[0][0], chain_detected
    print(_NOCHAIN)
    sys.exit()


class TcrInfo:
    """Everything ``import_tcr_info`` leaves in module globals, plus the device table handles."""

    def __init__(self, inputargs, quiet=False):
        self.chain, self.chain_detected = resolve_chain(inputargs)
        say = (lambda *a: None) if quiet else print
        say("Importing TCR", CHAINNAMS[self.chain], "gene sequences...")

        # tag-set / species fix-ups (decombine.py:640-675); inputargs is updated like the reference does
        if inputargs["tags"] == "extended" and inputargs["species"] == "mouse":
            say("Please note that there is currently no extended tag set for mouse TCR genes.\n"
                "     Decombinator will now switch the tag set in use from 'extended' to 'original'.")
            inputargs["tags"] = "original"
        if inputargs["tags"] == "extended" and self.chain in ("g", "d"):
            say("Please note that there is currently no extended tag set for gamma/delta TCR genes.\n"
                "     Decombinator will now switch the tag set in use from 'extended' to 'original'.")
            inputargs["tags"] = "original"
        if inputargs["tags"] == "extended":
            self.v_half_split, self.j_half_split = 10, 10
        elif inputargs["tags"] == "original":
            self.v_half_split, self.j_half_split = 10, 6
        else:
            print("Tag set unrecognised; should be either 'extended' or 'original' for human, or just 'original' "
                  "for mouse. \n     Please check tag set and species flag.")
            sys.exit()
        if inputargs["species"] not in ("human", "mouse"):
            print("Species not recognised. Please select either 'human' (default) or 'mouse'.")
            sys.exit()
        self.species, self.tags = inputargs["species"], inputargs["tags"]

        tagdir = inputargs.get("tagfastadir")
        for gene, split in (("v", self.v_half_split), ("j", self.j_half_split)):
            regions = parse_fasta_regions(read_tcr_text(self.species, self.tags, self.chain, gene, "fasta", tagdir))
            seqs, half1, half2, jumps = parse_tag_lines(
                read_tcr_text(self.species, self.tags, self.chain, gene, "tags", tagdir), split)
            setattr(self, gene + "_regions", regions)
            setattr(self, gene + "_seqs", seqs)
            setattr(self, "half1_" + gene + "_seqs", half1)
            setattr(self, "half2_" + gene + "_seqs", half2)
            setattr(self, "jump_to_end_v" if gene == "v" else "jump_to_start_j", jumps)
        self._tables = None

    def tables(self):
        """(V, J) dcb_tagset handles, built once."""
        if self._tables is None:
            self._tables = (
                _lib.TagTables(self.v_seqs, self.jump_to_end_v, self.v_regions, self.v_half_split, True),
                _lib.TagTables(self.j_seqs, self.jump_to_start_j, self.j_regions, self.j_half_split, False),
            )
        return self._tables


def load(species="human", tags="extended", chain="b", tagfastadir=None, quiet=True) -> TcrInfo:
    """Convenience wrapper for tests and benchmarks."""
    args = {"infile": "", "chain": chain, "tags": tags, "species": species, "tagfastadir": tagfastadir}
    return TcrInfo(args, quiet=quiet)
