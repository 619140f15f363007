function FI = interpolate(x, y, F, dx, dy)
% FI = interpolate(x, y, F, dx, dy)
% Drop-in for ray_trace_sw/interpolate.m (6x6 Lagrange stencil, bump 1e-13), evaluated by the
% LAGRANGE6 kernel of libswrt through the MEX gateway.  Shadow the original by putting this
% directory first on the path.
FI = swrt_mex('interpolate', x, y, F, dx, dy, 1e-13);
end
