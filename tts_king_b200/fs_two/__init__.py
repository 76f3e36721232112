"""Mirrors of the reference's ``fs_two`` call sites of the vocoder path (nothing else of fs_two)."""
