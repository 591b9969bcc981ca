"""B200-native residual-evaluation hot path of ImmersedBoundary.jl behind the package's own API.

Python host mirror of the Julia API on top of ``libibx.so`` (C ABI, ``include/ibx.h``); the Julia wrapper
a maintainer would ship is ``julia/ImmersedBoundaryB200.jl``.  No CPU fallback: importing this package
requires the built shared library, and every compute call requires a B200.
"""
from ._lib import IbxError  # noqa: F401
from .mesher import (Stereolitography, merge_points, refine_to_length, feature_regions, centers_and_normals,  # noqa: F401
                     Ball, Box, Line, Sphere, DistanceField, Mesh, get_cells)
from .domain import (Domain, Partition, Boundary, Surface, DeviceArray, Accumulator, Interpolator, context,  # noqa: F401
                     synchronize, launch_count, set_option, get_option, options, maximum, minimum, dot, at_owners, at_neighbors, at_faces, green_gauss,
                     unsigned_green_gauss, divergent, cell_gradient, face_distance, owner_distance, neighbor_distance,
                     face_gradient, JST_sensor, MUSCL, impose_bc, multigrid, volume_integral, surface_integral,
                     at_offset)
from .cfd import (Fluid, FlowBC, state2primitive, primitive2state, speed_of_sound, inviscid_fluxes,  # noqa: F401
                  dynamic_viscosity, heat_conductivity, viscous_fluxes, JST_sensor_3pt, shock_sensor, pressure_coefficient,
                  streamwise_direction, Reynolds_number, adjust_Reynolds,
                  residual_euler, residual_rans, step_euler, step_euler_sharded, ghost_update_euler, ghost_update_rans, residual_advection, euler_step_host, euler_step_host_begin,
                  euler_step_host_end, pinned_empty)
from .solver import FAS, march_euler, local_step_update, RK_STAGES, Multigrid, PIPreconditioner, hutchinson_trick, Linearization, linearize, proj_along, solve  # noqa: F401
from .vtk import export_vtk, vtk_grid_mesh, vtk_grid_stl  # noqa: F401
from . import synthetic, turbulence  # noqa: F401
