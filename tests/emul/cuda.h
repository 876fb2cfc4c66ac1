// CPU stand-in for <cuda.h> (driver API types the emulated sources mention) — TEST INFRASTRUCTURE.
#pragma once
struct CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
