function background_flow = grid_U(qk, K_d2, K2, kx_, ky_, shear_strength)
% background_flow = grid_U(qk, K_d2, K2, kx_, ky_, shear_strength)  -- qg_flow_ray_trace/grid_U.m
% The six spectral->grid transforms run on the device (cuFFT behind swrt_k2g).
if nargin < 6, shear_strength = 0; end
psik = -qk ./ (K_d2 + K2);
vk = 1i * kx_ .* psik;
uk = -1i * ky_ .* psik;
background_flow.u  = swrt_mex('k2g', uk) + shear_strength;
background_flow.v  = swrt_mex('k2g', vk);
background_flow.ux = swrt_mex('k2g', 1i * kx_ .* uk);
background_flow.uy = swrt_mex('k2g', 1i * ky_ .* uk);
background_flow.vx = swrt_mex('k2g', 1i * kx_ .* vk);
background_flow.vy = swrt_mex('k2g', 1i * ky_ .* vk);
end
