"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference UML hot path (uml_oracle.py), the harness that drives
the unmodified reference where it is available (ref_harness.py) and seeded synthetic bank
builders (synth.py).  Nothing under this directory is imported by the product package.
"""
