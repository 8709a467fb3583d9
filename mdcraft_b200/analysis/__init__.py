"""Analysis classes (mirror of ``mdcraft.analysis`` for the structure path only)."""
