classdef SpectralScheme < handle
    % scheme = SpectralScheme(L, nx, psi_field [, mode])   -- drop-in for SpectralScheme.m
    % mode 0 (default) = exact Fourier-series evaluation on the fp64 tensor pipe,
    % mode 1 = the reference's gridded 6x6 Lagrange evaluation.
    properties
        L, nx, mode, psik, eng
        U_field, GradU_field, psi_field          % the gridded planes the reference's class exposes (SpectralScheme.m:3,28-35)
    end
    methods
        function obj = SpectralScheme(L, nx, psi_field, mode)
            if nargin < 4, mode = 0; end
            obj.L = L; obj.nx = nx; obj.mode = mode;
            obj.psik = swrt_mex('g2k', psi_field);
            obj.eng = swrt_mex('create', nx, L, 1, 1, mode);
            swrt_mex('set_flow_spectral', obj.eng, 0, obj.psik);
            % the public gridded fields of the reference's class (SpectralScheme.m:16-35), transformed on the device; wavenumbers
            % in units of 2*pi/L, which is what the engine evaluates (the reference writes integer wavenumbers, i.e. L = 2*pi)
            kmax = nx/2 - 1;
            [kx_, ky_] = ndgrid(-kmax:kmax, 0:kmax);
            kap = 2*pi/L;
            ugk = -1i*kap*ky_.*obj.psik;  vgk = 1i*kap*kx_.*obj.psik;
            obj.psi_field = swrt_mex('k2g', obj.psik);
            obj.U_field.u = swrt_mex('k2g', ugk);
            obj.U_field.v = swrt_mex('k2g', vgk);
            obj.GradU_field.u_x = swrt_mex('k2g', 1i*kap*kx_.*ugk);
            obj.GradU_field.u_y = swrt_mex('k2g', 1i*kap*ky_.*ugk);
            obj.GradU_field.v_x = swrt_mex('k2g', 1i*kap*kx_.*vgk);
            obj.GradU_field.v_y = swrt_mex('k2g', 1i*kap*ky_.*vgk);
        end
        function delete(obj)
            if ~isempty(obj.eng), swrt_mex('destroy', obj.eng); end
        end
        function psi = streamfunction(obj, x, y, t) %#ok<INUSD>
            dx = obj.L / obj.nx;
            psi = reshape(swrt_mex('interpolate', x, y, swrt_mex('k2g', obj.psik), dx, dx), size(x));
        end
        function u = U(obj, x, t) %#ok<INUSD>
            xx = x(:, 1, :); yy = x(:, 2, :);
            [uu, vv] = swrt_mex('eval_at', obj.eng, 0, xx(:), yy(:));
            u = zeros(size(x));
            u(:, 1, :) = reshape(uu, size(xx));
            u(:, 2, :) = reshape(vv, size(xx));
        end
        function nablaU = grad_U(obj, x, t) %#ok<INUSD>
            xx = x(:, 1, :); yy = x(:, 2, :);
            [~, ~, nablaU.u_x, nablaU.u_y, nablaU.v_x, nablaU.v_y] = swrt_mex('eval_at', obj.eng, 0, xx(:), yy(:));
        end
        function nablaU_k = grad_U_times_k(obj, x, k, t)
            if nargin < 4, t = 0; end
            g = obj.grad_U(x, t);
            kk = k(:, 1, :); ll = k(:, 2, :);
            nablaU_k = zeros(size(k));
            nablaU_k(:, 1, :) = reshape(g.u_x .* kk(:) + g.v_x .* ll(:), size(kk));
            nablaU_k(:, 2, :) = reshape(g.u_y .* kk(:) + g.v_y .* ll(:), size(ll));
        end
    end
end
