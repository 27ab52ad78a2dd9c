"""ORACLE package — CPU checkers for the B200 hot path (test infrastructure only)."""
