"""Stub of matplotlib so that `import supervillain` works where matplotlib is absent (test infrastructure only)."""
