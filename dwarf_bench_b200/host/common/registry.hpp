// Forwarder: the reference keeps this part of the framework in common/registry.hpp; here it lives in one file.
#pragma once
#include "dwarf_framework.hpp"
