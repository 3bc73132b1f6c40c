"""Stand-in for ``lyft_dataset_sdk`` (pip, unpinned -- install_mods.sh:4 -- absent from this image).
TEST INFRASTRUCTURE ONLY; see ../README.md."""
