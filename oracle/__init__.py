"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference hot path (tb_oracle.c) plus the recipe that
builds the unmodified reference into oracle/_ref/. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product (tiebrush_b200/) never does."""
