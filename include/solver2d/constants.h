// solver2d-b200 — tuning constants of the s2World_Step path.
// Values are the contract of the drop-in boundary: they equal reference include/solver2d/constants.h:6-22 so that
// host C, the CUDA kernels and the oracle all bake in identical numbers.
#pragma once

#define s2_pi 3.14159265359f

// collision / solver tolerances (metres, radians)
#define s2_linearSlop 0.005f
#define s2_angularSlop (2.0f / 180.0f * s2_pi)
#define s2_speculativeDistance (4.0f * s2_linearSlop)
#define s2_aabbMargin 0.1f
#define s2_maxLinearCorrection 0.2f
#define s2_maxAngularCorrection (8.0f / 180.0f * s2_pi)
#define s2_huge (100000.0f)

// stabilisation
#define s2_baumgarte 0.2f
#define s2_maxBaumgarteVelocity 4.0f
#define s2_contactHertz 30.0f
#define s2_jointHertz 60.0f

// sleep thresholds (declared by the reference API, unused by any solver variant)
#define s2_timeToSleep 0.5f
#define s2_linearSleepTolerance 0.01f
#define s2_angularSleepTolerance (2.0f / 180.0f * s2_pi)

// capacities
#define s2_maxPolygonVertices 8
// The reference allows 32 worlds (constants.h:12). This library lifts the limit so that the 256-world batched
// configuration fits; ids stay int16 so the ABI is unchanged.
#define s2_maxWorlds 1024
