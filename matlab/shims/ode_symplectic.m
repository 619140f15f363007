function [x, k, t] = ode_symplectic(x0, k0, dt, T, f, gH, scheme, save_stride)
% [x, k, t] = ode_symplectic(x0, k0, dt, T, f, gH, scheme [, save_stride])  -- ode_symplectic.m
% All leapfrog steps between two saved rows run as ONE fused kernel launch with the packet state
% in registers.  save_stride (default 1 = the reference's behaviour) thins the stored history so
% that 64k-16M packets fit.
if nargin < 8, save_stride = 1; end
Nsteps = floor(T / dt);
rows = 0:save_stride:Nsteps - 1;
x = zeros([numel(rows), size(x0, [2, 3])]); k = x; t = zeros(numel(rows), 1);
x(1, :, :) = x0; k(1, :, :) = k0;
eng = swrt_mex('create', scheme.nx, scheme.L, f, gH, scheme.mode);
swrt_mex('set_flow_spectral', eng, 0, scheme.psik);
swrt_mex('set_packets', eng, squeeze(x0(1, 1, :)), squeeze(x0(1, 2, :)), squeeze(k0(1, 1, :)), squeeze(k0(1, 2, :)));
for r = 2:numel(rows)
    swrt_mex('step', eng, 0, dt, rows(r) - rows(r - 1));              % scheme 0 = leapfrog
    [px, py, pk, pl] = swrt_mex('get_packets', eng);
    x(r, 1, :) = px; x(r, 2, :) = py; k(r, 1, :) = pk; k(r, 2, :) = pl;
    t(r) = rows(r) * dt;
end
swrt_mex('destroy', eng);
end
