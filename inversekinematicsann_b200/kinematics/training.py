"""Training loop for the ANN solver -- the recipe of reference kinematics/ann.py:27-68.

Keras is not available offline, so the same recipe is restated with torch autograd (library GEMMs; training is
not the accelerated hot path -- inference of the trained weights runs through csrc/mlp*.cu):

* ``train_test_split(samples, features, test_size=0.33, random_state=42)`` (ann.py:29-30),
* ``StandardScaler`` fitted on the training split for inputs and labels (ann.py:32-36),
* ``Input(3)``, 12 x ``Dense(500, tanh)``, ``Dense(4)`` with Keras' defaults: Glorot-uniform kernels, zero biases
  (ann.py:46-56),
* ``Adam(learning_rate=1e-5)`` (Keras defaults beta 0.9/0.999, epsilon 1e-7), mean-squared error (ann.py:60),
* ``fit`` with validation on the held-out split, batch size 32 (Keras default; the reference leaves ``batch_size=64``
  commented out), reshuffled every epoch, ``EarlyStopping(monitor='val_loss', patience=12,
  restore_best_weights=True)`` (ann.py:58, 62-68).

``batch_size`` / ``learning_rate`` / ``final_learning_rate`` are exposed because the reference's 1e-5 x 32 schedule
needs hours per model; defaults stay the reference's.
"""
import numpy as np

HIDDEN_LAYERS, HIDDEN_UNITS = 12, 500  # net_shape of ann.py:48-50


def fit_training_data(samples, features):
    """(x_train, y_train, x_test, y_test, x_scaler, y_scaler) as reference ANN.__fit_trainig_data (ann.py:27-38)."""
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import StandardScaler
    x_train, x_test, y_train, y_test = train_test_split(samples, features, test_size=0.33, random_state=42)
    x_scaler, y_scaler = StandardScaler(), StandardScaler()
    x_train = x_scaler.fit_transform(x_train)
    x_test = x_scaler.transform(x_test)
    y_train = y_scaler.fit_transform(y_train)
    y_test = y_scaler.transform(y_test)
    return np.array(x_train), np.array(y_train), np.array(x_test), np.array(y_test), x_scaler, y_scaler


def train_dense_stack(x_train, y_train, x_test, y_test, epochs, batch_size=32, learning_rate=1.0e-5,
                      final_learning_rate=None, patience=12, hidden_layers=HIDDEN_LAYERS,
                      hidden_units=HIDDEN_UNITS, device=None, seed=None, verbose=True, allow_tf32=False):
    """Fit the tanh stack on already-scaled data; returns (kernels (in,out), biases, history dict)."""
    import torch
    if device is None:
        device = 'cuda' if torch.cuda.is_available() else 'cpu'
    if seed is not None:
        torch.manual_seed(seed)
    if str(device).startswith('cuda'):
        torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    dims = [x_train.shape[1]] + [hidden_units] * hidden_layers + [y_train.shape[1]]
    layers = []
    for i, (fan_in, fan_out) in enumerate(zip(dims[:-1], dims[1:])):
        dense = torch.nn.Linear(fan_in, fan_out)
        torch.nn.init.xavier_uniform_(dense.weight)  # Keras Dense default kernel_initializer
        torch.nn.init.zeros_(dense.bias)  # Keras Dense default bias_initializer
        layers.append(dense)
        if i < len(dims) - 2:
            layers.append(torch.nn.Tanh())
    net = torch.nn.Sequential(*layers).to(device)
    opt = torch.optim.Adam(net.parameters(), lr=learning_rate, betas=(0.9, 0.999), eps=1e-7)
    decay = 1.0
    if final_learning_rate is not None and epochs > 1:
        decay = (final_learning_rate / learning_rate) ** (1.0 / (epochs - 1))
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=decay)
    xt = torch.as_tensor(x_train, dtype=torch.float32, device=device)
    yt = torch.as_tensor(y_train, dtype=torch.float32, device=device)
    xv = torch.as_tensor(x_test, dtype=torch.float32, device=device)
    yv = torch.as_tensor(y_test, dtype=torch.float32, device=device)
    n = xt.shape[0]

    def validation_loss():
        net.eval()
        total = torch.zeros((), dtype=torch.float64, device=device)
        with torch.no_grad():
            for s in range(0, xv.shape[0], 65536):
                total += torch.nn.functional.mse_loss(net(xv[s:s + 65536]), yv[s:s + 65536], reduction='sum').double()
        net.train()
        return float(total) / max(1, yv.numel())

    history = {'loss': [], 'val_loss': []}
    best, best_state, waited = float('inf'), None, 0
    for epoch in range(epochs):
        order = torch.randperm(n, device=device)
        running = torch.zeros((), dtype=torch.float64, device=device)
        for s in range(0, n, batch_size):
            idx = order[s:s + batch_size]
            loss = torch.nn.functional.mse_loss(net(xt[idx]), yt[idx])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            running += loss.detach().double() * idx.numel()
        sched.step()
        val = validation_loss() if xv.shape[0] else float('nan')
        history['loss'].append(float(running) / n)
        history['val_loss'].append(val)
        if verbose:
            print(f'Epoch {epoch + 1}/{epochs} - loss: {history["loss"][-1]:.6f} - val_loss: {val:.6f}', flush=True)
        if val < best:  # EarlyStopping(monitor='val_loss', patience, restore_best_weights=True)
            best, waited = val, 0
            best_state = {k: v.detach().clone() for k, v in net.state_dict().items()}
        else:
            waited += 1
            if waited >= patience:
                break
    if best_state is not None:
        net.load_state_dict(best_state)
    dense = [m for m in net if isinstance(m, torch.nn.Linear)]
    kernels = [m.weight.detach().t().contiguous().cpu().numpy() for m in dense]  # Keras layout (in, out)
    biases = [m.bias.detach().cpu().numpy() for m in dense]
    history['best_val_loss'] = best
    return kernels, biases, history
