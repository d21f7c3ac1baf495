#pragma once
// Fills Registry::instance() with the dwarfs of this build (idempotent).  Reference: register_dwarfs.hpp.
void populate_registry();
