"""bench.py is what the driver runs unattended: every name a function of it (or of the tools it imports) reads as a global has to
exist in the module.  (A mis-placed edit once left `args` in a function that has no such variable: a NameError that only the
default run -- not the quick `--also none` runs used while developing -- would have met.)"""
import builtins
import importlib
import os
import symtable
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unresolved_globals(path, module):
    source = open(path).read()
    missing = []

    def walk(table):
        if table.get_type() == "function":
            for symbol in table.get_symbols():
                name = symbol.get_name()
                if symbol.is_global() and symbol.is_referenced() and not hasattr(module, name) and not hasattr(builtins, name):
                    missing.append((table.get_name(), name))
        for child in table.get_children():
            walk(child)

    walk(symtable.symtable(source, path, "exec"))
    return missing


@pytest.mark.parametrize("name", ["bench", "__graft_entry__"])
def test_every_global_a_function_reads_exists(name):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    module = importlib.import_module(name)
    assert unresolved_globals(os.path.join(ROOT, name + ".py"), module) == []
