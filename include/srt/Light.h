// srt/Light.h -- Light lives next to Material (both are plain data of the shaded variant).
#pragma once
#include "Material.h"
