"""alfred_margaret_b200 -- host-side mirror of alfred-margaret's Aho-Corasick API over libam_b200.so.

Module names follow the reference (Data.Text.AhoCorasick.{Automaton,Searcher,Replacer},
Data.Text.Utf8, Data.Text.CaseSensitivity).  All matching runs in the CUDA library; importing the
FFI without the built library raises ImportError (no CPU fallback).
"""
from . import automaton, case_sensitivity, replacer, searcher, sharded, splitter, synth, utf8  # noqa: F401
from .case_sensitivity import CaseSensitive, CaseSensitivity, IgnoreCase  # noqa: F401
