"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's (Illumina/Pisces) per-locus calling path, used as the parity checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs. The product (pisces_b200/) never imports this.
"""
