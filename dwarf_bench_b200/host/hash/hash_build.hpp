#pragma once
#include "join/b200_dwarfs.hpp"
