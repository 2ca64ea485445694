"""TEST INFRASTRUCTURE ONLY: CPU restatement ("oracle") of the reference's algorithm for the hot
path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here; the product package trinerflet_b200 never does."""
