"""Readers for the trajectory files MDCraft's OpenMM helpers write (``mdcraft.openmm.file``)."""
