"""TEST INFRASTRUCTURE ONLY: CPU oracle for the `blamm scan` hot path (see oracle.c / oracle.py)."""
