function FI = interpolate_par(x, y, F, dx, dy)
% Drop-in for interpolate_par.m (same stencil, bump 1e-10).
FI = swrt_mex('interpolate', x, y, F, dx, dy, 1e-10);
end
