# placeholder replaced below
