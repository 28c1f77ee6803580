"""Oracle = test infrastructure.  CPU restatements of the reference hot path used only as a checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
