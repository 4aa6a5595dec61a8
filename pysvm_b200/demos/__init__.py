"""Replays of the reference's demo scripts without the plotting."""
