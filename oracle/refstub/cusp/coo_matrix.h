// Stub: present only so that the reference header that includes it compiles; nothing from it is used.
#pragma once
#include <cusp/array1d.h>
