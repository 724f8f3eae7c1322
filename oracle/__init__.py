"""CPU oracles for the PDict matching path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package; nothing under biostrings_b200/ does.  See oracle/README.md.
"""
