"""ORACLE — test infrastructure only.  See oracle/orc_oracle.py and oracle/codecs.c."""
